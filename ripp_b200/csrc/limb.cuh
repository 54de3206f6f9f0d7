// 32-bit limb primitives with an explicit carry flag.
//
// Device build (sm_100a): one PTX instruction each (add.cc / madc.lo.cc ... -> IADD3.X / IMAD / IMAD.WIDE
// carry chains in SASS).  Host build (RIPP_HOSTSIM, used only by tests/hostsim to unit-test the
// *same* limb sequences on a CPU-only box): the carry flag is a thread-local variable.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RIPP_HD __host__ __device__ __forceinline__
// Functions at and above the Fq2-product level are real calls on the device: fully inlining a
// Miller loop (~10^5 instructions) makes ptxas run for hours and thrashes the instruction cache.
#define RIPP_FN __host__ __device__ __noinline__ inline
#else
#define RIPP_HD inline
#define RIPP_FN inline
#endif

#define RIPP_DEFCONST(name, n, ...)                \
  RIPP_HD uint32_t name(int i) {                   \
    const uint32_t a_[n] = {__VA_ARGS__};          \
    return a_[i];                                  \
  }

namespace ripp {
namespace limb {

#if defined(__CUDA_ARCH__)

RIPP_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
RIPP_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

#define RIPP_ASM3(fn, ins)                                                            \
  RIPP_HD void fn(uint32_t& r, uint32_t a, uint32_t b) {                              \
    asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                      \
  }
#define RIPP_ASM4(fn, ins)                                                            \
  RIPP_HD void fn(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) {                  \
    asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));          \
  }
RIPP_ASM3(add_cc, "add.cc.u32")
RIPP_ASM3(addc_cc, "addc.cc.u32")
RIPP_ASM3(addc, "addc.u32")
RIPP_ASM3(sub_cc, "sub.cc.u32")
RIPP_ASM3(subc_cc, "subc.cc.u32")
RIPP_ASM3(subc, "subc.u32")
RIPP_ASM4(mad_lo_cc, "mad.lo.cc.u32")
RIPP_ASM4(madc_lo_cc, "madc.lo.cc.u32")
RIPP_ASM4(mad_hi_cc, "mad.hi.cc.u32")
RIPP_ASM4(madc_hi_cc, "madc.hi.cc.u32")
RIPP_ASM4(madc_lo, "madc.lo.u32")
RIPP_ASM4(madc_hi, "madc.hi.u32")
#undef RIPP_ASM3
#undef RIPP_ASM4

// 32x32+64 -> 64 multiply-accumulate on an aligned register pair, with carry in/out.  Keeping the
// (lo, hi) halves in one 64-bit variable is what lets ptxas emit a single IMAD.WIDE.U32[.X]
// instead of IMAD + IMAD.HI + 2 IADD3.X when several products are in flight.
#define RIPP_MW(fn, ilo, ihi)                                                                         \
  RIPP_HD void fn(uint64_t& r, uint32_t a, uint32_t b, uint64_t c) {                                  \
    asm volatile("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %3;\n\t" ilo " l, %1, %2, l;\n\t" ihi         \
                 " h, %1, %2, h;\n\tmov.b64 %0, {l, h};\n\t}"                                          \
                 : "=l"(r)                                                                            \
                 : "r"(a), "r"(b), "l"(c));                                                           \
  }
RIPP_MW(mw_cc, "mad.lo.cc.u32", "madc.hi.cc.u32")
RIPP_MW(mwc_cc, "madc.lo.cc.u32", "madc.hi.cc.u32")
RIPP_MW(mwc, "madc.lo.cc.u32", "madc.hi.u32")
#undef RIPP_MW
RIPP_HD uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
// lo(x) += hi(y), carry out
RIPP_HD void add_lo_hi_cc(uint64_t& x, uint64_t y) {
  asm volatile("{\n\t.reg .u32 xl, xh, yl, yh;\n\tmov.b64 {xl, xh}, %0;\n\tmov.b64 {yl, yh}, %1;\n\t"
               "add.cc.u32 xl, xl, yh;\n\tmov.b64 %0, {xl, xh};\n\t}"
               : "+l"(x)
               : "l"(y));
}
// hi(x) += carry
RIPP_HD void addc_hi(uint64_t& x) {
  asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %0;\n\taddc.u32 xh, xh, 0;\n\tmov.b64 %0, {xl, xh};\n\t}"
               : "+l"(x));
}

#else  // host emulation of the same primitives

inline thread_local uint32_t cc_ = 0;

inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void add_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b;
  r = (uint32_t)t;
  cc_ = (uint32_t)(t >> 32);
}
inline void addc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + cc_;
  r = (uint32_t)t;
  cc_ = (uint32_t)(t >> 32);
}
inline void addc(uint32_t& r, uint32_t a, uint32_t b) { r = a + b + cc_; }
inline void sub_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b;
  r = (uint32_t)t;
  cc_ = (uint32_t)(t >> 32) & 1u;
}
inline void subc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - cc_;
  r = (uint32_t)t;
  cc_ = (uint32_t)(t >> 32) & 1u;
}
inline void subc(uint32_t& r, uint32_t a, uint32_t b) { r = a - b - cc_; }
inline void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { add_cc(r, mul_lo(a, b), c); }
inline void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { addc_cc(r, mul_lo(a, b), c); }
inline void mad_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { add_cc(r, mul_hi(a, b), c); }
inline void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { addc_cc(r, mul_hi(a, b), c); }
inline void madc_lo(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { addc(r, mul_lo(a, b), c); }
inline void madc_hi(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { addc(r, mul_hi(a, b), c); }

inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline void mw_cc(uint64_t& r, uint32_t a, uint32_t b, uint64_t c) {
  uint32_t l = (uint32_t)c, h = (uint32_t)(c >> 32);
  mad_lo_cc(l, a, b, l);
  madc_hi_cc(h, a, b, h);
  r = ((uint64_t)h << 32) | l;
}
inline void mwc_cc(uint64_t& r, uint32_t a, uint32_t b, uint64_t c) {
  uint32_t l = (uint32_t)c, h = (uint32_t)(c >> 32);
  madc_lo_cc(l, a, b, l);
  madc_hi_cc(h, a, b, h);
  r = ((uint64_t)h << 32) | l;
}
inline void mwc(uint64_t& r, uint32_t a, uint32_t b, uint64_t c) {
  uint32_t l = (uint32_t)c, h = (uint32_t)(c >> 32);
  madc_lo_cc(l, a, b, l);
  madc_hi(h, a, b, h);
  r = ((uint64_t)h << 32) | l;
}
inline void add_lo_hi_cc(uint64_t& x, uint64_t y) {
  uint32_t l = (uint32_t)x;
  add_cc(l, l, (uint32_t)(y >> 32));
  x = (x & 0xffffffff00000000ull) | l;
}
inline void addc_hi(uint64_t& x) {
  uint32_t h = (uint32_t)(x >> 32);
  addc(h, h, 0);
  x = (x & 0xffffffffull) | ((uint64_t)h << 32);
}

#endif

}  // namespace limb
}  // namespace ripp
