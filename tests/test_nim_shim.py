"""The Nim shim (numericalnim_b200/nim/b200rk.nim) cannot be compiled here (no Nim toolchain in the image), so its claim —
one {.importc, cdecl.} declaration per B200RK_API symbol of include/b200rk.h, same name, same number of parameters — is
checked textually: both files are parsed and compared."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "b200rk.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for proto in re.findall(r"B200RK_API\s+([^;{]+?);", h, flags=re.S):
        proto = " ".join(proto.split())
        m = re.match(r".+?\b(b200rk_\w+)\s*\((.*)\)$", proto)
        if not m:
            continue
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def nim_imports():
    src = open(os.path.join(ROOT, "numericalnim_b200", "nim", "b200rk.nim")).read()
    out = {}
    for m in re.finditer(r"^proc (b200rk_\w+)\((.*?)\)(?::\s*[\w\[\] ]+)?\s*\{\.importc, cdecl, dynlib: lib\.\}", src, flags=re.S | re.M):
        n = 0
        for group in [g for g in m.group(2).split(";") if g.strip()] if ";" in m.group(2) else [m.group(2)]:
            # "a, b: T, c: U" -> names before each ':' (a type applies to every name since the previous type)
            depth, cur, parts = 0, "", []
            for ch in group:
                depth += ch in "[(" 
                depth -= ch in "])"
                if ch == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += ch
            if cur.strip():
                parts.append(cur)
            n += len([p for p in parts if p.strip()])
        out[m.group(1)] = n
    return out


def test_every_header_symbol_has_an_importc_declaration_with_the_same_arity():
    hdr, nim = header_symbols(), nim_imports()
    assert len(hdr) >= 71
    assert sorted(set(hdr) - set(nim)) == [], "header symbols without a Nim declaration"
    assert sorted(set(nim) - set(hdr)) == [], "Nim declarations of symbols the header does not export"
    wrong = {k: (hdr[k], nim[k]) for k in hdr if hdr[k] != nim[k]}
    assert not wrong, wrong


def test_shim_mirrors_the_reference_interface_names():
    src = open(os.path.join(ROOT, "numericalnim_b200", "nim", "b200rk.nim")).read()
    for name in ("proc solveODE*(", "proc solveODEHost*(", "proc gpuIntegrator*(", "RK4_step*", "DOPRI54_step*", "TSIT54_step*", "VERN65_step*",
                 "proc hermiteSpline*(", "proc hermiteInterpolate*(", "proc newGpuVector*(", "proc newGpuSolver*("):
        assert name in src, name
    # exceptions never cross the ABI and ValueError is what the reference raises
    assert "newException(ValueError" in src and "except CatchableError" in src
