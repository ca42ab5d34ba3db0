// emul_lib_support.cpp — the pieces of the host-emulated library that are not transformed product sources: the state of
// the threaded launcher, and stand-ins for jit.cu (NVRTC + driver module API cannot be emulated; the PW_USER kernels are
// covered by tests/host_emul/emul_main.cpp instead). TEST INFRASTRUCTURE ONLY.
#include "internal.hpp"

EmulBlock* g_emul_block = nullptr;
std::binary_semaphore g_emul_device{1};
void emul_grid_sync() { if (g_emul_block) g_emul_block->block.arrive_and_wait(); }

static int unavailable(const b200rk_ctx* c) { return fail(c, B200RK_ECUDA, "right-hand sides from source need NVRTC and a GPU: not available under host emulation"); }
struct JitRhs { int unused; };
int jit_rhs_fn(double, const b200rk_vec* y, b200rk_vec*, void*) { return unavailable(y ? y->ctx : nullptr); }
void jit_describe(const JitRhs*, int* np, const b200rk_vec* const** vecs, const double** cs) { *np = 0; *vecs = nullptr; *cs = nullptr; }
int jit_slot_attempt(int) { return 0; }
int jit_slot_run(int) { return 0; }
int jit_launch(b200rk_ctx* c, JitRhs*, int, int, unsigned, void*, bool) { return unavailable(c); }
int jit_max_blocks_per_sm(b200rk_ctx* c, JitRhs*, int, int, int*) { return unavailable(c); }
int jit_prepare(b200rk_ctx* c, JitRhs*, int) { return unavailable(c); }
bool jit_is_stencil(const JitRhs*, int*, int*) { return false; }
int jit_launch_stencil_rk4(b200rk_ctx* c, JitRhs*, bool, double, double, const b200rk::L96Halo&, const b200rk_vec*, b200rk_vec*) { return unavailable(c); }
int jit_launch_rk4(b200rk_ctx* c, JitRhs*, bool, double, double, const b200rk_vec*, b200rk_vec*) { return unavailable(c); }

extern "C" {
int b200rk_jit_rhs_new(b200rk_ctx* c, const char*, int, const b200rk_vec* const*, int, const double*, b200rk_rhs_fn*, void**) { return unavailable(c); }
int b200rk_jit_stencil_rhs_new(b200rk_ctx* c, const char*, int, int, int, const b200rk_vec* const*, int, const double*, b200rk_rhs_fn*, void**) { return unavailable(c); }
int b200rk_jit_stencil_compile_only(const char*, int, int, int, int, int, size_t*, char*, size_t) { return unavailable(nullptr); }
int b200rk_jit_rhs_set_scalars(void*, int, const double*) { return unavailable(nullptr); }
int b200rk_jit_rhs_free(void*) { return B200RK_OK; }
int b200rk_jit_compile_only(const char*, int, int, int, void*, size_t, size_t*, char*, size_t) { return unavailable(nullptr); }
}
