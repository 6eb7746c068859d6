/* emu_cuda.h — host lock-step warp emulator (TEST/DEBUG ONLY, never part of the product).
 *
 * Lets tsid_control_b200/csrc/tsidb_kernels.cuh be compiled with g++ and executed on a
 * machine without a GPU: the 32 lanes of a warp run as 32 ucontext fibers that meet at every
 * warp collective (__shfl_sync, __shfl_xor_sync, __syncwarp).  It checks the kernel LOGIC
 * (indexing, phase hand-offs, pivot rules) against the oracle; it says nothing about
 * performance and is not a fallback — libtsidb.so does not contain it. */
#ifndef EMU_CUDA_H_
#define EMU_CUDA_H_
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define TSIDB_EMU 1
#define TSIDB_DEV static inline
#define TSIDB_DEVNI static
#define TSIDB_HD

namespace emu {
struct Warp {
  ucontext_t sched, ctx[32];
  char* stacks = nullptr;
  int cur = 0;
  bool done[32];
  uint64_t slot[32];
  int arrived = 0;
  unsigned gen = 0;
  long collectives = 0;
};
extern Warp W;
inline void yield() { swapcontext(&W.ctx[W.cur], &W.sched); }
inline void barrier() {
  unsigned my = W.gen;
  if (++W.arrived == 32) { W.arrived = 0; W.gen++; W.collectives++; }
  else while (W.gen == my) yield();
}
template <class T> inline T exchange(T x, int src) {
  static_assert(sizeof(T) <= 8, "");
  uint64_t raw = 0;
  memcpy(&raw, &x, sizeof(T));
  W.slot[W.cur] = raw;
  barrier();
  uint64_t got = W.slot[src & 31];
  barrier();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
}  // namespace emu

template <class T> inline T __shfl_sync(unsigned, T x, int src) { return emu::exchange(x, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T x, int m) { return emu::exchange(x, emu::W.cur ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T x, int d) { return emu::exchange(x, emu::W.cur + d > 31 ? emu::W.cur : emu::W.cur + d); }
inline void __syncwarp() { emu::barrier(); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  unsigned m = v;
  for (int o = 16; o > 0; o >>= 1) { unsigned t = emu::exchange(m, emu::W.cur ^ o); m = t < m ? t : m; }
  return m;
}
inline unsigned __ballot_sync(unsigned, bool p) {
  unsigned m = p ? (1u << emu::W.cur) : 0u;
  for (int o = 16; o > 0; o >>= 1) m |= emu::exchange(m, emu::W.cur ^ o);
  return m;
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline void sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
inline double2 __ldcs(const double2* p) { return *p; }
#endif
