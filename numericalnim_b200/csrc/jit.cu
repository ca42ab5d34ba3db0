// jit.cu — element-local right-hand sides compiled at run time INTO the fused kernels.
//
// The reference's extension point is the user's ODEProc closure (ode.nim:36): arbitrary host code returning
// dy/dt. A host closure cannot run inside a CUDA kernel, so the general path calls it once per stage
// (b200rk_rhs_fn) between stage kernels. For the large class of ELEMENT-LOCAL right-hand sides,
//      dydt[i] = expr(t, y[i], p0[i], .., p3[i], c0, .., c7),
// the caller can instead hand the library the expression as CUDA C++ source. NVRTC compiles kernels.cuh with
// that expression inlined as b200rk::user_rhs (PW_USER), giving the user's IVP exactly the kernels the built-in
// right-hand sides get: the whole-attempt kernel (5+np vector passes instead of ~57), the device-resident
// driver loop, the one-kernel RK4 step, plus a plain dydt = f(t, y) kernel for every other method. Compiled with
// -fmad=false: a*b+c in the expression is a multiply and an add, as in the reference's CPU arithmetic.
//
// NVRTC is bound with dlopen on first use and the driver's module API is reached through
// cudaGetDriverEntryPoint, so libb200rk.so links neither library and still loads on a machine without them.
#include "internal.hpp"
#include "stencil_attempt.cuh"

#include <cuda.h>
#include <nvrtc.h>  // types only

#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>

extern "C" const char b200rk_kernels_src[];  // kernels.cuh, embedded at build time (Makefile: kernels_src.cpp)
extern "C" const char b200rk_stencil_src[];  // stencil_attempt.cuh, likewise (stencil right-hand sides from source)

namespace {

// ---- NVRTC, bound lazily ------------------------------------------------------------------------------
struct NvrtcApi {
  void* handle = nullptr;
  std::string where;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  nvrtcResult (*Version)(int*, int*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
};
NvrtcApi g_nvrtc;
std::mutex g_jit_mutex;

bool file_exists(const std::string& p) {
  FILE* f = std::fopen(p.c_str(), "rb");
  if (f) std::fclose(f);
  return f != nullptr;
}

// The kernels use 256-bit global accesses (ld/st.global.v4.f64), which need the ptxas of the toolkit this library
// was built with (CUDA >= 12.9). A process may already hold an older libnvrtc.so.12 under the same SONAME (PyTorch
// bundles the one of its own toolkit), so the toolkit's copy is opened by PATH first — a second, private copy is
// fine, every entry point is taken from our own handle — and a candidate older than the build toolkit is only
// used when nothing newer can be found.
bool nvrtc_open(const std::string& name, NvrtcApi* a, int* version) {
  void* h = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  a->handle = h; a->where = name;
#define B200RK_BIND(field, sym) a->field = (decltype(a->field))dlsym(h, sym)
  B200RK_BIND(CreateProgram, "nvrtcCreateProgram");
  B200RK_BIND(DestroyProgram, "nvrtcDestroyProgram");
  B200RK_BIND(CompileProgram, "nvrtcCompileProgram");
  B200RK_BIND(GetProgramLogSize, "nvrtcGetProgramLogSize");
  B200RK_BIND(GetProgramLog, "nvrtcGetProgramLog");
  B200RK_BIND(GetCUBINSize, "nvrtcGetCUBINSize");
  B200RK_BIND(GetCUBIN, "nvrtcGetCUBIN");
  B200RK_BIND(AddNameExpression, "nvrtcAddNameExpression");
  B200RK_BIND(GetLoweredName, "nvrtcGetLoweredName");
  B200RK_BIND(GetErrorString, "nvrtcGetErrorString");
  B200RK_BIND(Version, "nvrtcVersion");
#undef B200RK_BIND
  if (!a->CreateProgram || !a->DestroyProgram || !a->CompileProgram || !a->GetProgramLogSize || !a->GetProgramLog || !a->GetCUBINSize ||
      !a->GetCUBIN || !a->AddNameExpression || !a->GetLoweredName || !a->GetErrorString || !a->Version)
    return false;
  int vmaj = 0, vmin = 0;
  if (a->Version(&vmaj, &vmin) != NVRTC_SUCCESS) return false;
  *version = vmaj * 1000 + vmin * 10;  // CUDART_VERSION encoding
  return true;
}

int nvrtc_bind(const b200rk_ctx* ctx) {
  if (g_nvrtc.handle) return B200RK_OK;
  std::vector<std::string> names;
  if (const char* p = getenv("B200RK_NVRTC_LIB")) names.push_back(p);
  const char* roots[] = {getenv("CUDA_HOME"), getenv("CUDA_PATH"), "/usr/local/cuda"};
  for (const char* root : roots)
    if (root) for (const char* n : {"/lib64/libnvrtc.so.12", "/lib64/libnvrtc.so", "/lib/libnvrtc.so.12"}) names.push_back(std::string(root) + n);
  for (const char* n : {"libnvrtc.so.12", "libnvrtc.so"}) names.push_back(n);
  NvrtcApi best;
  int best_version = 0;
  std::string tried;
  for (auto& n : names) {
    NvrtcApi a;
    int v = 0;
    if (!nvrtc_open(n, &a, &v)) { tried += n + "; "; continue; }
    if (v > best_version) { best = a; best_version = v; }
    if (v >= CUDART_VERSION / 10 * 10) break;  // as new as the toolkit libb200rk.so was built with
  }
  if (!best.handle) return fail(ctx, B200RK_ECUDA, "cannot load libnvrtc (set B200RK_NVRTC_LIB); tried: " + tried);
  g_nvrtc = best;
  return B200RK_OK;
}

// cooperative_groups.h lives in the toolkit's include directory
std::string cuda_include_dir() {
  std::vector<std::string> roots;
  if (const char* p = getenv("B200RK_CUDA_INCLUDE")) roots.push_back(p);
  const char* envs[] = {getenv("CUDA_HOME"), getenv("CUDA_PATH")};
  for (const char* e : envs) if (e) roots.push_back(std::string(e) + "/include");
  Dl_info info;
  if (g_nvrtc.CreateProgram && dladdr((void*)g_nvrtc.CreateProgram, &info) && info.dli_fname) {  // <root>/lib64/libnvrtc.so -> <root>/include
    std::string p = info.dli_fname;
    for (int up = 0; up < 2; ++up) { size_t k = p.find_last_of('/'); if (k == std::string::npos) break; p.erase(k); }
    roots.push_back(p + "/include");
    roots.push_back(p + "/targets/x86_64-linux/include");
  }
  roots.push_back("/usr/local/cuda/include");
  for (auto& r : roots) if (file_exists(r + "/cooperative_groups.h")) return r;
  return "";
}

// ---- kernels of one translation unit ------------------------------------------------------------------
enum { JB_RHS_W2 = 0, JB_RHS_W4_L2 = 1, JB_RK4_W4 = 2, JB_COUNT = 3 };              // base unit (pattern -1)
enum { JF_ATTEMPT_W4 = 0, JF_ATTEMPT_W2 = 1, JF_RUN_W4 = 2, JF_RUN_W2 = 3, JF_COUNT = 4 };  // one unit per FusedPattern
constexpr int kPatterns = 5;  // kernels.cuh: FusedPattern

// Stencil right-hand sides from source (stencil_attempt.cuh, B200RK_JIT_STENCIL): radii of the neighbourhood, 0 / 0 = element-local
struct StencilSpec {
  bool on = false;
  int rl = 0, rr = 0;
};
enum { JS_RHS = 0, JS_RK4 = 1 };   // stencil base unit
enum { JS_ATTEMPT = 0 };   // stencil unit of one FusedPattern

std::vector<std::string> name_expressions(int pattern, const StencilSpec& st = StencilSpec()) {
  const std::string T = std::to_string(kThreads);
  if (st.on) {
    if (pattern < 0) return {"b200rk::user_stencil_rhs_kernel<" + T + ">", "b200rk::ustencil_rk4_kernel<2, " + T + ">"};
    return {"b200rk::ustencil_attempt_kernel<" + std::to_string(pattern) + ", 2, " + T + ">"};
  }
  if (pattern < 0)
    return {"b200rk::user_rhs_kernel<2, 2, " + T + ", 0>", "b200rk::user_rhs_kernel<4, 2, " + T + ", 1>", "b200rk::user_rk4_kernel<4, " + T + ">"};
  const std::string P = std::to_string(pattern), K = std::to_string((int)PW_USER);
  return {"b200rk::fused_attempt_kernel<" + P + ", " + K + ", 4, " + T + ", 0>", "b200rk::fused_attempt_kernel<" + P + ", " + K + ", 2, " + T + ", 0>",
          "b200rk::fused_run_kernel<" + P + ", " + K + ", 4, " + T + ">", "b200rk::fused_run_kernel<" + P + ", " + K + ", 2, " + T + ">"};
}

struct Compiled {
  std::vector<char> cubin;
  std::vector<std::string> lowered;  // one per name expression
  std::string log;
};
std::map<std::string, Compiled> g_cubin_cache;  // identical (expr, np, nc, pattern) compile once per process

std::string make_stencil_source(const std::string& expr, int np, int nc, const StencilSpec& st) {
  std::string s = "#define B200RK_JIT_STENCIL 1\n#define B200RK_USER_NP " + std::to_string(np) + "\n#define B200RK_STENCIL_RL " + std::to_string(st.rl) +
                  "\n#define B200RK_STENCIL_RR " + std::to_string(st.rr) + "\nnamespace b200rk {\n";
  s += "__device__ __forceinline__ double user_stencil(double t, const double* b200rk_y_, const double* b200rk_p_, const double* b200rk_c_) {\n";
  s += "  (void)t; (void)b200rk_y_; (void)b200rk_p_; (void)b200rk_c_;\n";
  for (int j = 0; j < np; ++j) s += "  const double p" + std::to_string(j) + " = b200rk_p_[" + std::to_string(j) + "]; (void)p" + std::to_string(j) + ";\n";
  for (int j = 0; j < nc; ++j) s += "  const double c" + std::to_string(j) + " = b200rk_c_[" + std::to_string(j) + "]; (void)c" + std::to_string(j) + ";\n";
  // Y(d): the state at cyclic offset d; a d outside the declared radii is a compile error (negative array size)
  s += "#define Y(d) (b200rk_y_[(d) + 0 * (int)sizeof(char[((d) >= -" + std::to_string(st.rl) + " && (d) <= " + std::to_string(st.rr) + ") ? 1 : -1])])\n";
  s += "  return (double)(\n" + expr + "\n  );\n#undef Y\n}\n}  // namespace b200rk\n#include \"stencil_attempt.cuh\"\n";
  return s;
}

std::string make_source(const std::string& expr, int np, int nc) {
  std::string s = "#define B200RK_JIT 1\n#define B200RK_USER_NP " + std::to_string(np) + "\nnamespace b200rk {\n";
  s += "__device__ __forceinline__ double user_rhs(double t, double y, const double* b200rk_p_, const double* b200rk_c_) {\n";
  s += "  (void)t; (void)y; (void)b200rk_p_; (void)b200rk_c_;\n";
  for (int j = 0; j < np; ++j) s += "  const double p" + std::to_string(j) + " = b200rk_p_[" + std::to_string(j) + "]; (void)p" + std::to_string(j) + ";\n";
  for (int j = 0; j < nc; ++j) s += "  const double c" + std::to_string(j) + " = b200rk_c_[" + std::to_string(j) + "]; (void)c" + std::to_string(j) + ";\n";
  s += "  return (double)(\n" + expr + "\n  );\n}\n}  // namespace b200rk\n#include \"kernels.cuh\"\n";
  return s;
}

int validate(const b200rk_ctx* ctx, const char* expr, int np, int nc) {
  if (!expr || !*expr) return fail(ctx, B200RK_EINVAL, "jit rhs: empty expression");
  if (np < 0 || np > kMaxUserVecs) return fail(ctx, B200RK_EINVAL, "jit rhs: at most " + std::to_string(kMaxUserVecs) + " parameter vectors (p0..p3)");
  if (nc < 0 || nc > kMaxUserScalars) return fail(ctx, B200RK_EINVAL, "jit rhs: at most " + std::to_string(kMaxUserScalars) + " scalars (c0..c7)");
  const std::string e = expr;
  if (e.size() > 8192) return fail(ctx, B200RK_EINVAL, "jit rhs: expression longer than 8192 characters");
  if (e.find('#') != std::string::npos || e.find(';') != std::string::npos || e.find('{') != std::string::npos || e.find('}') != std::string::npos)
    return fail(ctx, B200RK_EINVAL, "jit rhs: a single C++ expression is expected (no '#', ';', '{', '}')");
  return B200RK_OK;
}

// ---- on-disk cache of compiled units ------------------------------------------------------------------
// $B200RK_JIT_CACHE (default ~/.cache/b200rk_jit; "0" or "off" disables). One file per (NVRTC version, kernels.cuh
// contents, np, nc, pattern, expression); the full key is stored in the file and compared on load.
uint64_t fnv1a(const char* p, size_t n, uint64_t h = 1469598103934665603ull) {
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
  return h;
}
std::string cache_dir() {
  const char* e = getenv("B200RK_JIT_CACHE");
  if (e) {
    if (!*e || !std::strcmp(e, "0") || !std::strcmp(e, "off")) return "";
    return e;
  }
  const char* home = getenv("HOME");
  return home && *home ? std::string(home) + "/.cache/b200rk_jit" : "";
}
std::string cache_path(const std::string& full_key) {
  const std::string dir = cache_dir();
  if (dir.empty()) return "";
  char name[32];
  std::snprintf(name, sizeof(name), "%016llx.b2rk", (unsigned long long)fnv1a(full_key.data(), full_key.size()));
  return dir + "/" + name;
}
void mkdirs(const std::string& dir) {
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') ::mkdir(dir.substr(0, i).c_str(), 0755);
}
bool cache_load(const std::string& full_key, Compiled* c) {
  const std::string path = cache_path(full_key);
  if (path.empty()) return false;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  auto rd = [&](void* p, size_t n) { return std::fread(p, 1, n, f) == n; };
  auto rd_str = [&](std::string* s2) { uint64_t n = 0; if (!rd(&n, 8) || n > (64u << 20)) return false; s2->resize(n); return n == 0 || rd(&(*s2)[0], n); };
  char magic[8];
  std::string key, blob;
  uint64_t nn = 0;
  bool ok = rd(magic, 8) && !std::memcmp(magic, "B2RKJIT1", 8) && rd_str(&key) && key == full_key && rd(&nn, 8) && nn <= 16;
  for (uint64_t i = 0; ok && i < nn; ++i) { std::string s2; ok = rd_str(&s2); if (ok) c->lowered.push_back(s2); }
  ok = ok && rd_str(&blob) && blob.size() > 4 && !std::memcmp(blob.data(), "\177ELF", 4);
  std::fclose(f);
  if (ok) c->cubin.assign(blob.begin(), blob.end());
  return ok;
}
void cache_store(const std::string& full_key, const Compiled& c) {
  const std::string path = cache_path(full_key);
  if (path.empty()) return;
  mkdirs(cache_dir());
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;
  auto wr = [&](const void* p, size_t n) { return std::fwrite(p, 1, n, f) == n; };
  auto wr_str = [&](const char* p, size_t n) { uint64_t n64 = n; return wr(&n64, 8) && (n == 0 || wr(p, n)); };
  uint64_t nn = c.lowered.size();
  bool ok = wr("B2RKJIT1", 8) && wr_str(full_key.data(), full_key.size()) && wr(&nn, 8);
  for (auto& s2 : c.lowered) ok = ok && wr_str(s2.data(), s2.size());
  ok = ok && wr_str(c.cubin.data(), c.cubin.size());
  ok = (std::fclose(f) == 0) && ok;
  if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

// NVRTC -> cubin for sm_100a. Compile errors are the caller's expression being wrong: B200RK_EINVAL + the log.
int compile_unit(const b200rk_ctx* ctx, const std::string& expr, int np, int nc, int pattern, const Compiled** out,
                 const StencilSpec& st = StencilSpec()) {
  std::lock_guard<std::mutex> lock(g_jit_mutex);
  const std::string key = std::to_string(np) + "|" + std::to_string(nc) + "|" + std::to_string(pattern) + "|" +
                          (st.on ? "stencil" + std::to_string(st.rl) + "," + std::to_string(st.rr) + "|" : "") + expr;
  auto hit = g_cubin_cache.find(key);
  if (hit != g_cubin_cache.end()) { *out = &hit->second; return B200RK_OK; }
  TRY(nvrtc_bind(ctx));
  int vmaj = 0, vmin = 0;
  if (g_nvrtc.Version) g_nvrtc.Version(&vmaj, &vmin);
  char stamp[64];
  std::snprintf(stamp, sizeof(stamp), "v1|nvrtc%d.%d|%016llx|", vmaj, vmin,
                (unsigned long long)fnv1a(b200rk_stencil_src, st.on ? std::strlen(b200rk_stencil_src) : 0,
                                          fnv1a(b200rk_kernels_src, std::strlen(b200rk_kernels_src))));
  const std::string full_key = stamp + key;
  {
    Compiled cached;
    if (cache_load(full_key, &cached)) { *out = &(g_cubin_cache[key] = std::move(cached)); return B200RK_OK; }
  }
  const std::string src = st.on ? make_stencil_source(expr, np, nc, st) : make_source(expr, np, nc);
  const char* hdr_src[] = {b200rk_kernels_src, b200rk_stencil_src};
  const char* hdr_name[] = {"kernels.cuh", "stencil_attempt.cuh"};
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "b200rk_user_rhs.cu", st.on ? 2 : 1, hdr_src, hdr_name);
  if (r != NVRTC_SUCCESS) return fail(ctx, B200RK_ECUDA, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
  const std::vector<std::string> names = name_expressions(pattern, st);
  for (auto& n : names) {
    r = g_nvrtc.AddNameExpression(prog, n.c_str());
    if (r != NVRTC_SUCCESS) { g_nvrtc.DestroyProgram(&prog); return fail(ctx, B200RK_ECUDA, "nvrtcAddNameExpression(" + n + "): " + g_nvrtc.GetErrorString(r)); }
  }
  std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false"};
  const std::string inc = cuda_include_dir();
  if (!inc.empty()) opts.push_back("--include-path=" + inc);
  std::vector<const char*> optv;
  for (auto& o : opts) optv.push_back(o.c_str());
  r = g_nvrtc.CompileProgram(prog, (int)optv.size(), optv.data());
  Compiled c;
  size_t log_n = 0;
  if (g_nvrtc.GetProgramLogSize(prog, &log_n) == NVRTC_SUCCESS && log_n > 1) {
    c.log.resize(log_n);
    g_nvrtc.GetProgramLog(prog, &c.log[0]);
    while (!c.log.empty() && (c.log.back() == '\0' || c.log.back() == '\n')) c.log.pop_back();
  }
  if (r != NVRTC_SUCCESS) {
    g_nvrtc.DestroyProgram(&prog);
    const bool user_error = (r == NVRTC_ERROR_COMPILATION);
    return fail(ctx, user_error ? B200RK_EINVAL : B200RK_ECUDA,
                std::string("jit rhs: ") + g_nvrtc.GetErrorString(r) + " for expression `" + expr + "`" +
                    (inc.empty() ? " (CUDA include directory with cooperative_groups.h not found: set B200RK_CUDA_INCLUDE)" : "") + "\n" + c.log);
  }
  size_t nb = 0;
  r = g_nvrtc.GetCUBINSize(prog, &nb);
  if (r == NVRTC_SUCCESS && nb) { c.cubin.resize(nb); r = g_nvrtc.GetCUBIN(prog, c.cubin.data()); }
  if (r != NVRTC_SUCCESS || !nb) { g_nvrtc.DestroyProgram(&prog); return fail(ctx, B200RK_ECUDA, "nvrtcGetCUBIN failed"); }
  for (auto& n : names) {
    const char* low = nullptr;
    r = g_nvrtc.GetLoweredName(prog, n.c_str(), &low);
    if (r != NVRTC_SUCCESS || !low) { g_nvrtc.DestroyProgram(&prog); return fail(ctx, B200RK_ECUDA, "nvrtcGetLoweredName(" + n + ") failed"); }
    c.lowered.push_back(low);
  }
  g_nvrtc.DestroyProgram(&prog);
  cache_store(full_key, c);
  *out = &(g_cubin_cache[key] = std::move(c));
  return B200RK_OK;
}

// ---- driver module API through the runtime ------------------------------------------------------------
struct DrvApi {
  bool ready = false;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*LaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};
DrvApi g_drv;

int drv_bind(const b200rk_ctx* c) {
  if (g_drv.ready) return B200RK_OK;
  DrvApi a;
  auto get = [&](const char* sym, void** fn) -> bool {
    cudaDriverEntryPointQueryResult st = cudaDriverEntryPointSymbolNotFound;
    cudaError_t e = cudaGetDriverEntryPoint(sym, fn, cudaEnableDefault, &st);
    if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fn) { cudaGetLastError(); return false; }
    return true;
  };
  bool ok = get("cuModuleLoadData", (void**)&a.ModuleLoadData) && get("cuModuleUnload", (void**)&a.ModuleUnload) &&
            get("cuModuleGetFunction", (void**)&a.ModuleGetFunction) && get("cuLaunchKernel", (void**)&a.LaunchKernel) &&
            get("cuLaunchCooperativeKernel", (void**)&a.LaunchCooperativeKernel) &&
            get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&a.OccupancyMaxActiveBlocksPerMultiprocessor) &&
            get("cuGetErrorString", (void**)&a.GetErrorString);
  if (!ok) return fail(c, B200RK_ECUDA, "jit rhs: the CUDA driver's module API is not reachable (cudaGetDriverEntryPoint)");
  a.ready = true;
  g_drv = a;
  return B200RK_OK;
}

std::string drv_error(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUDA driver error " + std::to_string((int)r);
}
#define DRV_TRY(ctx, expr)                                                                         \
  do {                                                                                             \
    CUresult _r = (expr);                                                                          \
    if (_r != CUDA_SUCCESS) return fail(ctx, B200RK_ECUDA, std::string(#expr) + ": " + drv_error(_r)); \
  } while (0)

struct JitModule {
  CUmodule mod = nullptr;
  CUfunction fn[4] = {nullptr, nullptr, nullptr, nullptr};
};

}  // namespace

struct JitRhs {
  b200rk_ctx* ctx = nullptr;
  std::string expr;
  int np = 0, nc = 0;
  StencilSpec stencil;   // on: dydt[i] depends on y[i - rl .. i + rr] (cyclic); off: element-local
  const b200rk_vec* vecs[kMaxUserVecs] = {nullptr};
  double cs[kMaxUserScalars] = {0};
  JitModule base, pat[kPatterns];
};

static int ensure_module(b200rk_ctx* c, JitRhs* j, int pattern, JitModule** out) {
  if (pattern >= kPatterns) return fail(c, B200RK_EINVAL, "jit rhs: bad pattern");
  JitModule* m = pattern < 0 ? &j->base : &j->pat[pattern];
  if (!m->mod) {
    const Compiled* cc = nullptr;
    TRY(compile_unit(c, j->expr, j->np, j->nc, pattern, &cc, j->stencil));
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaFree(nullptr));  // the runtime's primary context is current on this thread from here on
    TRY(drv_bind(c));
    CUmodule mod = nullptr;
    DRV_TRY(c, g_drv.ModuleLoadData(&mod, cc->cubin.data()));
    for (size_t i = 0; i < cc->lowered.size(); ++i) {
      CUresult r = g_drv.ModuleGetFunction(&m->fn[i], mod, cc->lowered[i].c_str());
      if (r != CUDA_SUCCESS) { g_drv.ModuleUnload(mod); return fail(c, B200RK_ECUDA, "cuModuleGetFunction(" + cc->lowered[i] + "): " + drv_error(r)); }
    }
    m->mod = mod;
  }
  *out = m;
  return B200RK_OK;
}

// ---- used by executor.cu ------------------------------------------------------------------------------
int jit_launch(b200rk_ctx* c, JitRhs* j, int pattern, int slot, unsigned grid, void* arg_block, bool cooperative) {
  JitModule* m = nullptr;
  TRY(ensure_module(c, j, pattern, &m));
  void* params[] = {arg_block};
  if (cooperative) DRV_TRY(c, g_drv.LaunchCooperativeKernel(m->fn[slot], grid, 1, 1, kThreads, 1, 1, 0, (CUstream)c->stream, params));
  else DRV_TRY(c, g_drv.LaunchKernel(m->fn[slot], grid, 1, 1, kThreads, 1, 1, 0, (CUstream)c->stream, params, nullptr));
  return B200RK_OK;
}

// Compile (or fetch from the caches) and load everything a solver over this right-hand side will launch — the plain
// dy = f(t, y) unit and, when the method has a fused form, its pattern unit — BEFORE the first reducing launch: a first-use
// NVRTC compile takes seconds, and a rank that compiles lazily inside its first attempt keeps its peers' kernels spinning
// on the mailbox for that long.
int jit_prepare(b200rk_ctx* c, JitRhs* j, int pattern) {
  JitModule* m = nullptr;
  TRY(ensure_module(c, j, -1, &m));
  if (pattern >= 0) TRY(ensure_module(c, j, pattern, &m));
  return B200RK_OK;
}

int jit_max_blocks_per_sm(b200rk_ctx* c, JitRhs* j, int pattern, int slot, int* per_sm) {
  JitModule* m = nullptr;
  TRY(ensure_module(c, j, pattern, &m));
  DRV_TRY(c, g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(per_sm, m->fn[slot], kThreads, 0));
  return B200RK_OK;
}

int jit_slot_attempt(int w) { return w == 4 ? JF_ATTEMPT_W4 : JF_ATTEMPT_W2; }
int jit_slot_run(int w) { return w == 4 ? JF_RUN_W4 : JF_RUN_W2; }

void jit_describe(const JitRhs* j, int* np, const b200rk_vec* const** vecs, const double** cs) {
  *np = j->np; *vecs = j->vecs; *cs = j->cs;
}

static int fill_user_args(b200rk_ctx* c, const JitRhs* j, const b200rk_vec* y, UserRhsArgs* a) {
  std::memset(a, 0, sizeof(*a));
  for (int i = 0; i < j->np; ++i) { TRY(check_same(c, y, j->vecs[i])); a->p[i] = j->vecs[i]->d; }
  for (int i = 0; i < j->nc; ++i) a->cs[i] = j->cs[i];
  a->y = y->d; a->n = y->n_local; a->tsign = 1.0; a->rsign = 1.0;
  return B200RK_OK;
}

// Stencil right-hand side from source, plain evaluation: one launch; sharded, the rl elements before and the rr after the
// shard travel first (one grouped ncclSend/ncclRecv with the ring neighbours on the context stream, like the built-in Lorenz-96).
static int jit_stencil_eval(b200rk_ctx* c, JitRhs* j, double t, const b200rk_vec* y, b200rk_vec* dydt) {
  const size_t n = y->n_local;
  if (dydt->d == y->d) return fail(c, B200RK_EINVAL, "stencil rhs: the output must not alias the input");
  if (y->n_global < (size_t)(j->stencil.rl + j->stencil.rr + 1)) return fail(c, B200RK_EINVAL, "stencil rhs: the vector is shorter than the stencil");
  UStencilRhsArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int i = 0; i < j->np; ++i) { TRY(check_same(c, y, j->vecs[i])); a.p[i] = j->vecs[i]->d; }
  for (int i = 0; i < j->nc; ++i) a.cs[i] = j->cs[i];
  a.y = y->d; a.t = t; a.out = dydt->d; a.n = n;
  if (c->world > 1) {
    const int rl = j->stencil.rl, rr = j->stencil.rr;
    for (int r = 0; r < c->world; ++r) {   // the same verdict on every rank
      size_t o_ = 0, l_ = 0;
      shard_range(y->n_global, r, c->world, &o_, &l_);
      if (l_ < (size_t)std::max(rl, rr)) return fail(c, B200RK_EINVAL, "stencil rhs: every shard must hold at least max(radius_left, radius_right) elements");
    }
    if (!c->d_halo_stencil) CUDA_TRY(c, cudaMalloc(&c->d_halo_stencil, 2 * kMaxStencilRadius * sizeof(double)));
    double *hl = c->d_halo_stencil, *hr = c->d_halo_stencil + kMaxStencilRadius;
    const int left = (c->rank + c->world - 1) % c->world, right = (c->rank + 1) % c->world;
    NCCL_TRY(c, g_nccl.GroupStart());
    if (rr) NCCL_TRY(c, g_nccl.Send(y->d, rr, ncclDouble, left, c->comm, c->stream));            // my head is the left neighbour's right halo
    if (rl) NCCL_TRY(c, g_nccl.Send(y->d + n - rl, rl, ncclDouble, right, c->comm, c->stream));   // my tail is the right neighbour's left halo
    if (rr) NCCL_TRY(c, g_nccl.Recv(hr, rr, ncclDouble, right, c->comm, c->stream));
    if (rl) NCCL_TRY(c, g_nccl.Recv(hl, rl, ncclDouble, left, c->comm, c->stream));
    NCCL_TRY(c, g_nccl.GroupEnd());
    c->collectives++;
    a.left = hl; a.right = hr;
  }
  if (!n) return B200RK_OK;
  ProfScope ps(c, B200RK_K_RHS, 8.0 * double(n) * (2 + j->np));
  const unsigned grid = (unsigned)std::min<size_t>((n + kThreads * 4 - 1) / (kThreads * 4), (size_t)c->sm_count * 16);
  return jit_launch(c, j, -1, JS_RHS, grid, &a, false);
}

// One RK4 step of a stencil right-hand side from source in one kernel (stencil_attempt.cuh: ustencil_rk4_kernel).
int jit_launch_stencil_rk4(b200rk_ctx* c, JitRhs* j, bool negate, double t, double dt, const L96Halo& halo, const b200rk_vec* y, b200rk_vec* y_new) {
  const size_t n = y->n_local;
  if (!n) return B200RK_OK;
  UStencilRk4Args a;
  std::memset(&a, 0, sizeof(a));
  for (int i = 0; i < j->np; ++i) { TRY(check_same(c, y, j->vecs[i])); a.p[i] = j->vecs[i]->d; }
  for (int i = 0; i < j->nc; ++i) a.cs[i] = j->cs[i];
  a.y = y->d; a.ynew = y_new->d; a.n = n;
  a.t = t; a.tsign = negate ? -1.0 : 1.0; a.rsign = negate ? -1.0 : 1.0;
  a.hdt = 0.5 * dt; a.dt = dt; a.c6 = dt / 6.0;   // same host scalars as launch_fused_rk4 / launch_rk4_final
  a.halo = halo;
  const int OUT = 4 * kThreads - stencil_halo(j->stencil.rl, 5) - stencil_halo(j->stencil.rr, 5);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(n) * (2 + j->np));
  return jit_launch(c, j, -1, JS_RK4, (unsigned)((n + OUT - 1) / OUT), &a, false);
}

bool jit_is_stencil(const JitRhs* j, int* rl, int* rr) {
  if (rl) *rl = j->stencil.rl;
  if (rr) *rr = j->stencil.rr;
  return j->stencil.on;
}

// The ODEProc itself: dydt = f(t, y), one launch (what the stage / RHS / finish pipeline calls per stage).
int jit_rhs_fn(double t, const b200rk_vec* y, b200rk_vec* dydt, void* user) {
  JitRhs* j = static_cast<JitRhs*>(user);
  b200rk_ctx* c = j->ctx;
  TRY(check_same(c, y, dydt));
  const size_t n = y->n_local;
  if (n == 0) return B200RK_OK;
  if (j->stencil.on) return jit_stencil_eval(c, j, t, y, dydt);
  UserRhsArgs a;
  TRY(fill_user_args(c, j, y, &a));
  a.t = t; a.out = dydt->d;
  ProfScope ps(c, B200RK_K_RHS, 8.0 * double(n) * (2 + j->np));
  if (l2_on(c, n) && dydt->d != y->d)  // same hand-off policy as the built-in right-hand sides (launch.cu: launch_ewise_t)
    return jit_launch(c, j, -1, JB_RHS_W4_L2, grid_for(c, n / 4, kThreads * 2), &a, false);
  return jit_launch(c, j, -1, JB_RHS_W2, grid_for(c, n / 2, kThreads * 2), &a, false);
}

// One RK4 step in one kernel (ode.nim:180-189): reads y (+ parameters), writes yNew.
int jit_launch_rk4(b200rk_ctx* c, JitRhs* j, bool negate, double t, double dt, const b200rk_vec* y, b200rk_vec* y_new) {
  const size_t n = y->n_local;
  if (!n) return B200RK_OK;
  UserRhsArgs a;
  TRY(fill_user_args(c, j, y, &a));
  a.t = t; a.tsign = negate ? -1.0 : 1.0; a.rsign = negate ? -1.0 : 1.0;
  a.hdt = 0.5 * dt; a.dt = dt; a.c6 = dt / 6.0; a.out = y_new->d;
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(n) * (2 + j->np));
  return jit_launch(c, j, -1, JB_RK4_W4, grid_for(c, n / 4, kThreads, c->ctas_per_sm), &a, false);
}

extern "C" {

int b200rk_jit_rhs_new(b200rk_ctx* c, const char* expr, int n_vec, const b200rk_vec* const* vecs, int n_scalar,
                       const double* scalars, b200rk_rhs_fn* fn, void** user) {
  if (!c || !fn || !user) return fail(c, B200RK_EINVAL, "null argument");
  TRY(validate(c, expr, n_vec, n_scalar));
  if ((n_vec > 0 && !vecs) || (n_scalar > 0 && !scalars)) return fail(c, B200RK_EINVAL, "jit rhs: null parameter array");
  JitRhs* j = new JitRhs;
  j->ctx = c; j->expr = expr; j->np = n_vec; j->nc = n_scalar;
  for (int i = 0; i < n_vec; ++i) {
    if (!vecs[i] || vecs[i]->ctx != c) { delete j; return fail(c, B200RK_EINVAL, "jit rhs: parameter vector belongs to another context"); }
    if (i && vecs[i]->n_global != vecs[0]->n_global) { delete j; return fail(c, B200RK_EINVAL, "Vectors must have the same size."); }
    j->vecs[i] = vecs[i];
  }
  for (int i = 0; i < n_scalar; ++i) j->cs[i] = scalars[i];
  // compile + load the base unit now, so a wrong expression is reported here (B200RK_EINVAL + compiler log)
  JitModule* m = nullptr;
  int rc = ensure_module(c, j, -1, &m);
  if (rc != B200RK_OK) { delete j; return rc; }
  *fn = &jit_rhs_fn;
  *user = j;
  return B200RK_OK;
}

int b200rk_jit_stencil_rhs_new(b200rk_ctx* c, const char* expr, int radius_left, int radius_right, int n_vec, const b200rk_vec* const* vecs,
                               int n_scalar, const double* scalars, b200rk_rhs_fn* fn, void** user) {
  if (!c || !fn || !user) return fail(c, B200RK_EINVAL, "null argument");
  TRY(validate(c, expr, n_vec, n_scalar));
  if (radius_left < 0 || radius_right < 0 || radius_left > kMaxStencilRadius || radius_right > kMaxStencilRadius)
    return fail(c, B200RK_EINVAL, "stencil rhs: radii must be in 0.." + std::to_string(kMaxStencilRadius));
  if ((n_vec > 0 && !vecs) || (n_scalar > 0 && !scalars)) return fail(c, B200RK_EINVAL, "jit rhs: null parameter array");
  JitRhs* j = new JitRhs;
  j->ctx = c; j->expr = expr; j->np = n_vec; j->nc = n_scalar;
  j->stencil.on = true; j->stencil.rl = radius_left; j->stencil.rr = radius_right;
  for (int i = 0; i < n_vec; ++i) {
    if (!vecs[i] || vecs[i]->ctx != c) { delete j; return fail(c, B200RK_EINVAL, "jit rhs: parameter vector belongs to another context"); }
    if (i && vecs[i]->n_global != vecs[0]->n_global) { delete j; return fail(c, B200RK_EINVAL, "Vectors must have the same size."); }
    j->vecs[i] = vecs[i];
  }
  for (int i = 0; i < n_scalar; ++i) j->cs[i] = scalars[i];
  JitModule* m = nullptr;
  int rc = ensure_module(c, j, -1, &m);   // a wrong expression (or a Y(d) outside the radii) is reported here with the compiler log
  if (rc != B200RK_OK) { delete j; return rc; }
  *fn = &jit_rhs_fn;
  *user = j;
  return B200RK_OK;
}

int b200rk_jit_rhs_set_scalars(void* user, int n_scalar, const double* scalars) {
  JitRhs* j = static_cast<JitRhs*>(user);
  if (!j) return fail(nullptr, B200RK_EINVAL, "null argument");
  if (n_scalar != j->nc || (n_scalar > 0 && !scalars)) return fail(j->ctx, B200RK_EINVAL, "jit rhs: scalar count differs from the compiled expression's");
  for (int i = 0; i < n_scalar; ++i) j->cs[i] = scalars[i];  // kernel arguments: takes effect at the next launch
  return B200RK_OK;
}

int b200rk_jit_rhs_free(void* user) {
  JitRhs* j = static_cast<JitRhs*>(user);
  if (!j) return B200RK_OK;
  if (j->ctx) cudaStreamSynchronize(j->ctx->stream);
  if (g_drv.ready) {
    if (j->base.mod) g_drv.ModuleUnload(j->base.mod);
    for (auto& m : j->pat) if (m.mod) g_drv.ModuleUnload(m.mod);
  }
  delete j;
  return B200RK_OK;
}

// Host only (no device, no context): NVRTC-compile one translation unit for sm_100a. pattern -1 = the plain
// dydt / RK4 kernels, 0..4 = the fused attempt + device-loop kernels of that FusedPattern.
int b200rk_jit_compile_only(const char* expr, int n_vec, int n_scalar, int pattern, void* cubin_out, size_t cubin_cap,
                            size_t* cubin_bytes, char* log, size_t log_cap) {
  TRY(validate(nullptr, expr, n_vec, n_scalar));
  if (pattern < -1 || pattern >= kPatterns) return fail(nullptr, B200RK_EINVAL, "jit rhs: pattern must be in -1..4");
  const Compiled* cc = nullptr;
  int rc = compile_unit(nullptr, expr, n_vec, n_scalar, pattern, &cc);
  if (rc != B200RK_OK) {
    if (log && log_cap) { std::strncpy(log, thread_error().c_str(), log_cap - 1); log[log_cap - 1] = '\0'; }
    return rc;
  }
  if (cubin_bytes) *cubin_bytes = cc->cubin.size();
  if (cubin_out && cubin_cap >= cc->cubin.size()) std::memcpy(cubin_out, cc->cubin.data(), cc->cubin.size());
  if (log && log_cap) {
    std::string l = cc->log;
    for (auto& n : cc->lowered) l += "\nkernel " + n;
    std::strncpy(log, l.c_str(), log_cap - 1);
    log[log_cap - 1] = '\0';
  }
  return B200RK_OK;
}

// Host only: the stencil units (pattern -1 = the dydt kernel, 0..4 = the whole-attempt kernel of that FusedPattern).
int b200rk_jit_stencil_compile_only(const char* expr, int radius_left, int radius_right, int n_vec, int n_scalar, int pattern, size_t* cubin_bytes,
                                    char* log, size_t log_cap) {
  TRY(validate(nullptr, expr, n_vec, n_scalar));
  if (pattern < -1 || pattern >= kPatterns) return fail(nullptr, B200RK_EINVAL, "jit rhs: pattern must be in -1..4");
  if (radius_left < 0 || radius_right < 0 || radius_left > kMaxStencilRadius || radius_right > kMaxStencilRadius)
    return fail(nullptr, B200RK_EINVAL, "stencil rhs: radii must be in 0.." + std::to_string(kMaxStencilRadius));
  StencilSpec st;
  st.on = true; st.rl = radius_left; st.rr = radius_right;
  const Compiled* cc = nullptr;
  int rc = compile_unit(nullptr, expr, n_vec, n_scalar, pattern, &cc, st);
  if (rc != B200RK_OK) {
    if (log && log_cap) { std::strncpy(log, thread_error().c_str(), log_cap - 1); log[log_cap - 1] = '\0'; }
    return rc;
  }
  if (cubin_bytes) *cubin_bytes = cc->cubin.size();
  if (log && log_cap) {
    std::string l = cc->log;
    for (auto& n : cc->lowered) l += "\nkernel " + n;
    std::strncpy(log, l.c_str(), log_cap - 1);
    log[log_cap - 1] = '\0';
  }
  return B200RK_OK;
}

}  // extern "C"
