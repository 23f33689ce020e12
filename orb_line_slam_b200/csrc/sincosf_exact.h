// Bit-exact restatement of glibc's sinf / cosf (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized-routines
// algorithm glibc has shipped since 2.28) for |x| < 120, so that the per-keypoint trig of rBRIEF
// (a = cosf(angle * pi/180), b = sinf(...), reference src/ORBextractor.cc:113-115) can run ON THE DEVICE and still equal the
// host libm result the CPU reference produces.
//
// Range reduction and both polynomials are evaluated in double precision and rounded once to float; that makes the float
// result insensitive to FMA contraction: the same source with and without fused multiply-adds reproduces libm's sinf and
// cosf on ALL 1 087 373 313 floats in [0, 6.5] (exhaustive host check: tests/test_trig_exact.py, run against the libm
// of the machine the tests run on; the device version is swept against host libm in tests/test_gpu_orb.py).
// The coefficients are the published polynomial of that algorithm (also readable in libm.so's __sincosf_table).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define OLF_TRIG_HD __host__ __device__ __forceinline__
#else
#define OLF_TRIG_HD inline
#endif

namespace olf {
namespace trig {

struct Poly { double c0, c1, c2, c3, c4, s1, s2, s3; };
// cosine polynomial negated in the second set (quadrants 2, 3); the sine sign comes from sign[n & 3]
OLF_TRIG_HD Poly poly_set(int neg) {
    Poly p;
    p.c0 = neg ? -0x1p0 : 0x1p0;
    p.c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    p.c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
    p.c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    p.c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    p.s1 = -0x1.555545995a603p-3; p.s2 = 0x1.1107605230bc4p-7; p.s3 = -0x1.994eb3774cf24p-13;
    return p;
}
#if defined(__CUDA_ARCH__)
OLF_TRIG_HD double tmul(double a, double b) { return __dmul_rn(a, b); }
OLF_TRIG_HD double tadd(double a, double b) { return __dadd_rn(a, b); }
OLF_TRIG_HD uint32_t fbits(float f) { return __float_as_uint(f); }
#else
OLF_TRIG_HD double tmul(double a, double b) { return a * b; }
OLF_TRIG_HD double tadd(double a, double b) { return a + b; }
OLF_TRIG_HD uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#endif
OLF_TRIG_HD uint32_t abstop12(float x) { return (fbits(x) >> 20) & 0x7ff; }

OLF_TRIG_HD float sincos_poly(double x, double x2, const Poly& p, int n) {
    if ((n & 1) == 0) {
        const double x3 = tmul(x, x2);
        const double s1 = tadd(p.s2, tmul(x2, p.s3));
        const double x7 = tmul(x3, x2);
        const double s = tadd(x, tmul(x3, p.s1));
        return (float)tadd(s, tmul(x7, s1));
    }
    const double x4 = tmul(x2, x2);
    const double c2 = tadd(p.c3, tmul(x2, p.c4));
    const double c1 = tadd(p.c0, tmul(x2, p.c1));
    const double x6 = tmul(x4, x2);
    const double c = tadd(c1, tmul(x4, p.c2));
    return (float)tadd(c, tmul(x6, c2));
}
// x - n * pi/2 with n = round(x * 2/pi); valid for |x| < 120
OLF_TRIG_HD double reduce_fast(double x, int* np) {
    const double r = tmul(x, 0x1.45F306DC9C883p+23);                 // 2/pi * 2^24
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return tadd(x, -tmul((double)n, 0x1.921FB54442D18p0));
}
OLF_TRIG_HD double quadrant_sign(int n) { return ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0; }     // sign[4] = {1, -1, -1, 1}

// the caller guarantees |y| < 120 (angles of rBRIEF are in [0, 2 pi])
OLF_TRIG_HD float sinf_exact(float y) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincos_poly(x, tmul(x, x), poly_set(0), 0);
    }
    int n;
    x = reduce_fast(x, &n);
    return sincos_poly(tmul(x, quadrant_sign(n)), tmul(x, x), poly_set(n & 2), n);
}
OLF_TRIG_HD float cosf_exact(float y) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, tmul(x, x), poly_set(0), 1);
    }
    int n;
    x = reduce_fast(x, &n);
    return sincos_poly(tmul(x, quadrant_sign(n)), tmul(x, x), poly_set(n & 2), n ^ 1);
}

}  // namespace trig
}  // namespace olf
