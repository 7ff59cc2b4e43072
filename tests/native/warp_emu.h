// warp_emu.h — TEST INFRASTRUCTURE. A 32-lane warp on the CPU: every lane is a fiber (ucontext) running the SAME device function; the warp
// intrinsics (__ballot_sync, __shfl_sync, __any_sync, __syncwarp, ...) are rendezvous points at which the fibers are switched round-robin, so the
// lanes advance in lockstep exactly where the GPU requires them to. A collective that not all 32 lanes reach is a deadlock on the GPU too; here
// it is reported. Lets tests/native/codec_warp_host.cc run cassandra_b200/csrc/lz4.cuh and snappy.cuh (the warp-per-chunk compressors K5 uses)
// under g++ and compare them with the oracle and the golden vectors without a GPU.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace warp_emu {
enum { LANES = 32, STACK = 256 * 1024 };
struct Warp {
    ucontext_t main_ctx, ctx[LANES]; std::vector<char> stacks; int cur = 0; bool done[LANES]; int ndone = 0;
    uint64_t slot[2][LANES]; int arrived = 0; unsigned phase = 0; int kind[LANES];
    std::function<void(int)> body;
};
static thread_local Warp* W = nullptr;
inline int lane() { return W->cur; }
inline void next_lane() {                 // switch to the next lane that has not finished; back to main when all have
    Warp* w = W; int from = w->cur;
    for (int k = 1; k <= LANES; k++) { int l = (from + k) % LANES; if (!w->done[l]) { w->cur = l; if (l != from) swapcontext(&w->ctx[from], &w->ctx[l]); return; } }
}
// deposit a value, let every other live lane deposit too, then read the whole vector
inline const uint64_t* collective(uint64_t v, int kind_id) {
    Warp* w = W; const unsigned ph = w->phase & 1; const int me = w->cur;
    w->slot[ph][me] = v; w->kind[me] = kind_id; w->arrived++;
    const int live = LANES - w->ndone;
    if (w->arrived < live) { next_lane(); }              // others still have to arrive: run them; we come back when the last one switched on
    else {                                                // last arrival completes the collective: every live lane must be in the SAME one
        for (int l = 0; l < LANES; l++) if (!w->done[l] && w->kind[l] != kind_id) { fprintf(stderr, "warp_emu: divergent collective (lane %d kind %d vs lane %d kind %d)\n", me, kind_id, l, w->kind[l]); abort(); }
        w->arrived = 0; w->phase++;
    }
    // when a waiting lane runs again every live lane has deposited: in round-robin order the last arriver flips the phase and goes on, the others
    // are resumed one after the other as it reaches ITS next rendezvous (which uses the other slot buffer) or finishes
    return w->slot[ph];
}
static void trampoline() { Warp* w = W; int me = w->cur; w->body(me); w->done[me] = true; w->ndone++;
    if (w->arrived && w->arrived >= LANES - w->ndone) { fprintf(stderr, "warp_emu: lane %d finished while others wait in a collective\n", me); abort(); }
    if (w->ndone == LANES) { setcontext(&w->main_ctx); }
    for (int k = 1; k <= LANES; k++) { int l = (me + k) % LANES; if (!w->done[l]) { w->cur = l; setcontext(&w->ctx[l]); } }
}
inline void run(std::function<void(int)> body) {
    Warp w; w.body = body; w.stacks.resize((size_t)LANES * STACK); memset(w.done, 0, sizeof(w.done)); memset(w.kind, 0, sizeof(w.kind));
    W = &w;
    for (int l = 0; l < LANES; l++) { getcontext(&w.ctx[l]); w.ctx[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * STACK; w.ctx[l].uc_stack.ss_size = STACK; w.ctx[l].uc_link = nullptr; makecontext(&w.ctx[l], trampoline, 0); }
    w.cur = 0; swapcontext(&w.main_ctx, &w.ctx[0]);
    W = nullptr;
}
} // namespace warp_emu

#define FULL_MASK_EMU 0xFFFFFFFFu
static inline unsigned __ballot_sync(unsigned, int pred) { const uint64_t* s = warp_emu::collective(pred ? 1 : 0, 1); unsigned r = 0; for (int l = 0; l < 32; l++) if (s[l]) r |= 1u << l; return r; }   // (a lane that returned right after its last collective still counts in it)
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xFFFFFFFFu; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); const uint64_t* s = warp_emu::collective(x, 2); T r; memcpy(&r, &s[src & 31], sizeof(T)); return r; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); const uint64_t* s = warp_emu::collective(x, 3); T r; memcpy(&r, &s[(warp_emu::lane() ^ d) & 31], sizeof(T)); return r; }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { warp_emu::collective(0, 4); }
static inline void __threadfence_block() {}
