// VQIP algebra for the sm_100a kernels (reference: wsimod/core/core.py).
//
// A VQIP is V = P+1 fp64 lanes held in registers: x[0] = volume, x[1..P] pollutants in
// constants.POLLUTANTS order.  P is a template parameter so every loop unrolls and the
// lanes are statically indexed (a dynamically indexed array would be demoted to local
// memory).  The pollutant configuration CFG supplies the additive mask and the planes
// of the named pollutants either as compile-time constants (default WSIMOD pollutant
// set) or as run-time values read from the descriptor.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define WSI_HD __host__ __device__ __forceinline__
#define WSI_HDN __host__ __device__ __noinline__
#else
#define WSI_HD inline
#define WSI_HDN
#endif

#include <math.h>

namespace wsi {

template <int P>
struct VQ {
  double x[P + 1];
};

// --- pollutant configurations ---------------------------------------------------------
// Default WSIMOD pollutants (constants.py:21-50): do, org-phosphorus, phosphate, ammonia,
// solids, bod, cod, ph, temperature, nitrate, nitrite, org-nitrogen; do/ph/temperature
// are non-additive.
struct CfgDefault12 {
  static constexpr int P = 12;
  static constexpr bool kStatic = true;
  static constexpr bool kLeanKernel = true;  // also build the schedule kernel without the rare node classes
  WSI_HD uint32_t add_mask() const { return 0x1CFCu; }  // planes 2-7 and 10-12, see below
  WSI_HD int no3() const { return 10; }
  WSI_HD int no2() const { return 11; }
  WSI_HD int nh4() const { return 4; }
  WSI_HD int po4() const { return 3; }
  WSI_HD int orgp() const { return 2; }
  WSI_HD int orgn() const { return 12; }
  WSI_HD int solids() const { return 5; }
  WSI_HD int temp() const { return 9; }
};
// additive planes of the default set: org-phosphorus(2) phosphate(3) ammonia(4) solids(5)
// bod(6) cod(7) nitrate(10) nitrite(11) org-nitrogen(12) -> bits
static_assert(((1u << 2) | (1u << 3) | (1u << 4) | (1u << 5) | (1u << 6) | (1u << 7) | (1u << 10) |
               (1u << 11) | (1u << 12)) == 0x1CFCu,
              "default additive mask");

template <int P_>
struct CfgRuntime {
  static constexpr int P = P_;
  static constexpr bool kStatic = false;
  static constexpr bool kLeanKernel = false;
  uint32_t mask;
  int i_no3, i_no2, i_nh4, i_po4, i_orgp, i_orgn, i_solids, i_temp;
  WSI_HD uint32_t add_mask() const { return mask; }
  WSI_HD int no3() const { return i_no3; }
  WSI_HD int no2() const { return i_no2; }
  WSI_HD int nh4() const { return i_nh4; }
  WSI_HD int po4() const { return i_po4; }
  WSI_HD int orgp() const { return i_orgp; }
  WSI_HD int orgn() const { return i_orgn; }
  WSI_HD int solids() const { return i_solids; }
  WSI_HD int temp() const { return i_temp; }
};

// Run-time named planes but a compile-time additive mask: the mask is what every VQIP operation
// branches on, so fixing it specialises the hot code; instantiated for the Oxford demo's pollutant
// set (12 pollutants, temperature the only non-additive one: mask 0x1FFC minus its bit).
template <int P_, uint32_t MASK>
struct CfgStaticMask : CfgRuntime<P_> {
  static constexpr bool kLeanKernel = true;
  WSI_HD uint32_t add_mask() const { return MASK; }
};

// --- lane access by (possibly run-time) plane index -------------------------------------
// With a compile-time index the unrolled select chain folds to a single register.
template <int P>
WSI_HD double vget(const VQ<P>& v, int i) {
  double r = 0.0;
#pragma unroll
  for (int p = 0; p <= P; p++)
    if (p == i) r = v.x[p];
  return r;
}
template <int P>
WSI_HD void vset(VQ<P>& v, int i, double val) {
#pragma unroll
  for (int p = 0; p <= P; p++)
    if (p == i) v.x[p] = val;
}
template <int P>
WSI_HD void vadd(VQ<P>& v, int i, double val) {
#pragma unroll
  for (int p = 0; p <= P; p++)
    if (p == i) v.x[p] += val;
}
template <int P>
WSI_HD void vsub(VQ<P>& v, int i, double val) {
#pragma unroll
  for (int p = 0; p <= P; p++)
    if (p == i) v.x[p] -= val;
}
template <int P>
WSI_HD void vmul(VQ<P>& v, int i, double val) {
#pragma unroll
  for (int p = 0; p <= P; p++)
    if (p == i) v.x[p] *= val;
}

// --- algebra (core.py) ------------------------------------------------------------------
template <int P>
WSI_HD VQ<P> vq_zero() {  // empty_vqip core.py:32
  VQ<P> v;
#pragma unroll
  for (int p = 0; p <= P; p++) v.x[p] = 0.0;
  return v;
}

WSI_HD bool bit(uint32_t m, int p) { return (m >> p) & 1u; }

// num / den for den > 0, returning a zero numerator as it is ((+-0)/x == +-0).  The fp64 division
// sequence leaves its inline fast path (for a ~150-instruction subroutine) when the quotient is
// zero, and the compiler turns `num == 0 ? num : num / den` into divide-then-select, so the
// divide itself is given a harmless numerator.  The empty asm keeps the optimiser from folding
// that numerator back into `num`.
WSI_HD double div_z(double num, double den) {
  double n = num == 0.0 ? 1.0 : num;
#if defined(__CUDA_ARCH__)
  asm("" : "+d"(n));
#endif
  const double q = n / den;
  return num == 0.0 ? num : q;
}

// sum_vqip core.py:83-116: volume and additive add, non-additive volume-weighted blend
// (t2 term first), first operand survives when the total volume is zero.
template <int P>
WSI_HD VQ<P> vq_sum(uint32_t add, const VQ<P>& a, const VQ<P>& b) {
  VQ<P> t;
  t.x[0] = a.x[0] + b.x[0];
  const bool wet = t.x[0] > 0;
  const double den = wet ? t.x[0] : 1.0;  // a zero denominator would take the slow division path
#pragma unroll
  for (int p = 1; p <= P; p++) {
    if (bit(add, p))
      t.x[p] = a.x[p] + b.x[p];
    else
    {
      // 0/den is +-0 with the numerator's sign: return it directly -- a zero numerator would
      // send the fp64 division down its slow path
      const double num = b.x[p] * b.x[0] + a.x[p] * a.x[0];
      t.x[p] = wet ? div_z(num, den) : a.x[p];
    }
  }
  return t;
}

// extract_vqip core.py:150-178
template <int P>
WSI_HD VQ<P> vq_extract(uint32_t add, const VQ<P>& a, const VQ<P>& b) {
  VQ<P> t;
  t.x[0] = a.x[0] - b.x[0];
#pragma unroll
  for (int p = 1; p <= P; p++) t.x[p] = bit(add, p) ? a.x[p] - b.x[p] : a.x[p];
  return t;
}

// v_change_vqip core.py:251-284: `volume *= v/volume`, or plain assignment when dry
template <int P>
WSI_HD VQ<P> vq_vchange(uint32_t add, const VQ<P>& a, double v) {
  VQ<P> t = a;
  if (a.x[0] > 0) {
    const double ratio = div_z(v, a.x[0]);
    t.x[0] = a.x[0] * ratio;
#pragma unroll
    for (int p = 1; p <= P; p++)
      if (bit(add, p)) t.x[p] = a.x[p] * ratio;
  } else {
    t.x[0] = v;
  }
  return t;
}
// the volume lane of v_change_vqip alone
WSI_HD double vol_vchange(double vol, double v) { return vol > 0 ? vol * div_z(v, vol) : v; }

// a VQIP that carries only a volume: v_change_vqip(empty_vqip(), v)
template <int P>
WSI_HD VQ<P> vq_volume(double v) {
  VQ<P> t = vq_zero<P>();
  t.x[0] = v;
  return t;
}

// A parameter plane (never written by a kernel): read through the non-coherent path, which lets
// the compiler move the load across the kernel's stores -- every plane of a group hangs off the same
// base pointer, so it cannot prove that for an ordinary load.
WSI_HD double ldp(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// min(a, b) with Python semantics (`a if a <= b`... `min(a, b)` returns a unless b < a)
WSI_HD double pymin(double a, double b) { return b < a ? b : a; }
WSI_HD double pymax(double a, double b) { return b > a ? b : a; }

}  // namespace wsi
