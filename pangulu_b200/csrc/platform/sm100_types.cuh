// sm100_types.cuh -- value-type arithmetic and device-side block descriptors for the
// sm_100a PanguLU platform module.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pangulu_sm100_abi.h"

namespace psm {

typedef uint32_t u32;
typedef uint16_t u16;

// ---- value arithmetic -----------------------------------------------------------------------
// calculate_type is double / float / {re,im} pairs (pangulu_sm100_abi.h).  Everything the kernels
// need: zero, a*b, x -= a*b, x / pivot, the reference's pivot clamp.
template <class T> struct is_complex { static const bool value = false; };
template <> struct is_complex<pangulu_sm100_c64> { static const bool value = true; };
template <> struct is_complex<pangulu_sm100_c32> { static const bool value = true; };

__host__ __device__ inline double vzero(double*) { return 0.0; }
__host__ __device__ inline float vzero(float*) { return 0.0f; }
__host__ __device__ inline pangulu_sm100_c64 vzero(pangulu_sm100_c64*) { pangulu_sm100_c64 z; z.re = 0; z.im = 0; return z; }
__host__ __device__ inline pangulu_sm100_c32 vzero(pangulu_sm100_c32*) { pangulu_sm100_c32 z; z.re = 0; z.im = 0; return z; }
template <class T> __host__ __device__ inline T zero() { return vzero((T*)0); }

__host__ __device__ inline bool is_zero(double a) { return a == 0.0; }
__host__ __device__ inline bool is_zero(float a) { return a == 0.0f; }
__host__ __device__ inline bool is_zero(pangulu_sm100_c64 a) { return a.re == 0.0 && a.im == 0.0; }
__host__ __device__ inline bool is_zero(pangulu_sm100_c32 a) { return a.re == 0.0f && a.im == 0.0f; }

// x -= a*b
__device__ inline void fnma(double& x, double a, double b) { x = fma(-a, b, x); }
__device__ inline void fnma(float& x, float a, float b) { x = fmaf(-a, b, x); }
__device__ inline void fnma(pangulu_sm100_c64& x, pangulu_sm100_c64 a, pangulu_sm100_c64 b)
{
    x.re = fma(-a.re, b.re, x.re); x.re = fma(a.im, b.im, x.re);
    x.im = fma(-a.re, b.im, x.im); x.im = fma(-a.im, b.re, x.im);
}
__device__ inline void fnma(pangulu_sm100_c32& x, pangulu_sm100_c32 a, pangulu_sm100_c32 b)
{
    x.re = fmaf(-a.re, b.re, x.re); x.re = fmaf(a.im, b.im, x.re);
    x.im = fmaf(-a.re, b.im, x.im); x.im = fmaf(-a.im, b.re, x.im);
}
// x += a*b
__device__ inline void fma_acc(double& x, double a, double b) { x = fma(a, b, x); }
__device__ inline void fma_acc(float& x, float a, float b) { x = fmaf(a, b, x); }
__device__ inline void fma_acc(pangulu_sm100_c64& x, pangulu_sm100_c64 a, pangulu_sm100_c64 b)
{
    x.re = fma(a.re, b.re, x.re); x.re = fma(-a.im, b.im, x.re);
    x.im = fma(a.re, b.im, x.im); x.im = fma(a.im, b.re, x.im);
}
__device__ inline void fma_acc(pangulu_sm100_c32& x, pangulu_sm100_c32 a, pangulu_sm100_c32 b)
{
    x.re = fmaf(a.re, b.re, x.re); x.re = fmaf(-a.im, b.im, x.re);
    x.im = fmaf(a.re, b.im, x.im); x.im = fmaf(a.im, b.re, x.im);
}
__device__ inline void vsub(double& x, double a) { x -= a; }
__device__ inline void vsub(float& x, float a) { x -= a; }
__device__ inline void vsub(pangulu_sm100_c64& x, pangulu_sm100_c64 a) { x.re -= a.re; x.im -= a.im; }
__device__ inline void vsub(pangulu_sm100_c32& x, pangulu_sm100_c32 a) { x.re -= a.re; x.im -= a.im; }

__device__ inline double vdiv(double a, double b) { return a / b; }
__device__ inline float vdiv(float a, float b) { return a / b; }
// complex division as gcc's C99 _Complex '/' evaluates it in the reference CPU code would be Smith-like;
// the straightforward formula below agrees to rounding on the well-scaled pivots of static-pivot LU.
__device__ inline pangulu_sm100_c64 vdiv(pangulu_sm100_c64 a, pangulu_sm100_c64 b)
{
    double d = b.re * b.re + b.im * b.im;
    pangulu_sm100_c64 r;
    r.re = (a.re * b.re + a.im * b.im) / d;
    r.im = (a.im * b.re - a.re * b.im) / d;
    return r;
}
__device__ inline pangulu_sm100_c32 vdiv(pangulu_sm100_c32 a, pangulu_sm100_c32 b)
{
    float d = b.re * b.re + b.im * b.im;
    pangulu_sm100_c32 r;
    r.re = (a.re * b.re + a.im * b.im) / d;
    r.im = (a.im * b.re - a.re * b.im) / d;
    return r;
}

__device__ inline double vmul(double a, double b) { return a * b; }
__device__ inline float vmul(float a, float b) { return a * b; }
__device__ inline pangulu_sm100_c64 vmul(pangulu_sm100_c64 a, pangulu_sm100_c64 b)
{
    pangulu_sm100_c64 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
__device__ inline pangulu_sm100_c32 vmul(pangulu_sm100_c32 a, pangulu_sm100_c32 b)
{
    pangulu_sm100_c32 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
__device__ inline double vrcp(double p) { return 1.0 / p; }
__device__ inline float vrcp(float p) { return 1.0f / p; }
__device__ inline pangulu_sm100_c64 vrcp(pangulu_sm100_c64 p) { pangulu_sm100_c64 o; o.re = 1.0; o.im = 0.0; return vdiv(o, p); }
__device__ inline pangulu_sm100_c32 vrcp(pangulu_sm100_c32 p) { pangulu_sm100_c32 o; o.re = 1.0f; o.im = 0.0f; return vdiv(o, p); }

// Pivot clamp of the CPU platform (pangulu_platform_0100000.c:80-84,153-157):
//   if (fabs(pivot) < PANGULU_TOL) pivot = PANGULU_TOL;
// fabs() on a C99 complex converts to double first, i.e. looks at the REAL part only, and the
// replacement value is the real number 1e-16.
__device__ inline double clamp_pivot(double p) { return fabs(p) < PANGULU_TOL ? PANGULU_TOL : p; }
__device__ inline float clamp_pivot(float p) { return fabs((double)p) < PANGULU_TOL ? (float)PANGULU_TOL : p; }
__device__ inline pangulu_sm100_c64 clamp_pivot(pangulu_sm100_c64 p)
{
    if (fabs(p.re) < PANGULU_TOL) { p.re = PANGULU_TOL; p.im = 0.0; }
    return p;
}
__device__ inline pangulu_sm100_c32 clamp_pivot(pangulu_sm100_c32 p)
{
    if (fabs((double)p.re) < PANGULU_TOL) { p.re = (float)PANGULU_TOL; p.im = 0.0f; }
    return p;
}

// ---- device-side descriptors ------------------------------------------------------------------
// One sparse block as the kernels see it.  ptr/idx/val are the CSC trio (CSR trio for the upper half
// of a diagonal block).  rptr/ridx/rmap are the CSR mirror of a strictly-lower block (else null).
template <class T>
struct Blk
{
    T* val;
    const u32* ptr;
    const u16* idx;
    const u32* rptr;
    const u16* ridx;
    const u32* rmap;
};

// GETRF: lower = strict lower CSC, upper = upper incl. diagonal CSR (diagonal first in each row)
// [c0, c1): the index range of the block to factor (an independent group of the diagonal block; the whole block: 0, nb)
template <class T> struct GetrfTask { Blk<T> lower, upper; int c0, c1, grp, pad; };
// TSTRF: b = strictly-lower block (uses its CSR mirror), diag = UPPER half of the diagonal block
// GESSM: b = upper block (CSC), diag = LOWER half of the diagonal block
// dense: index of the block's compact dense copy (sm100_compact.cuh) or -1
// kind: PANGULU_TASK_TSTRF / _GESSM in the low byte, bits 8..11: groups of the diagonal block the solve depends on
template <class T> struct PanelTask { Blk<T> b, diag; int kind; int dense; };
// SSSSM: destination (c = the block, or the LOWER half when the destination is a diagonal block;
// u = UPPER half or null) and the range of (A,B) operand pairs applied to it
// read-only CSC operand of an SSSSM pair
template <class T> struct BlkRO { const T* val; const u32* ptr; const u16* idx; };
template <class T> struct SsssmDest { Blk<T> c; Blk<T> u; u32 pair_begin, pair_end; u32 nnz_c, flags; };
template <class T> struct SsssmPair { BlkRO<T> a, b; };

}  // namespace psm
