// biquad_math.cuh — RBJ coefficient formulas of BiQuadFilterNode.UpdateCoefficients (Nodes/BiQuadFilterNode.cs:149-258) with a
// bit-exact restatement of glibc sinf/cosf.  Include only from translation units compiled with --fmad=false.
#pragma once
#include "gac_kernels.h"

namespace gac {

// glibc's sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h — the ARM optimized-routines algorithm that .NET's
// MathF.Sin/Cos reach through the platform libm on Linux): double-precision pi/2 reduction + double polynomial,
// rounded once to float.  Restated here so that the RBJ coefficients (BiQuadFilterNode.cs:151-153) come out
// bit-identical to the CPU oracle; verified exhaustively on the host for every float in [1e-5, 3.2].
// Valid for 0 <= x < 120 (w0 = 2*pi*f/fs lies in (0, pi]).
__device__ __forceinline__ void sincosf_libm(float y, float* sn, float* cs) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  double x = (double)y;
  int n = 0;
  double sgn = 1.0;
  bool neg = false;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
  if (top >= 0x3f4) {  // |y| >= pi/4  (abstop12(pio4) = 0x3f4)
    double r = x * hpi_inv;
    n = ((int)r + 0x800000) >> 24;
    x = x - (double)n * hpi;
    sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    neg = (n & 2) != 0;
  } else if (top < 0x398) {  // |y| < 2^-12
    *sn = y;
    *cs = 1.0f;
    return;
  }
  const double x2 = x * x;
  const double xs = x * sgn;
  // sine polynomial on (xs, x2), cosine polynomial on (x2) with the sign of table 1 when n & 2
  double sinp, cosp;
  {
    double x3 = xs * x2;
    double S1 = s2 + x2 * s3;
    double x7 = x3 * x2;
    double sv = xs + x3 * s1;
    sinp = sv + x7 * S1;
  }
  {
    double C0 = neg ? -c0 : c0, C1 = neg ? -c1 : c1, C2 = neg ? -c2 : c2, C3 = neg ? -c3 : c3, C4 = neg ? -c4 : c4;
    double x4 = x2 * x2;
    double cc2 = C3 + x2 * C4;
    double cc1 = C0 + x2 * C1;
    double x6 = x4 * x2;
    double cv = cc1 + x4 * C2;
    cosp = cv + x6 * cc2;
  }
  // sinf uses poly(n), cosf uses poly(n ^ 1): even -> sine polynomial, odd -> cosine polynomial
  if ((n & 1) == 0) {
    *sn = (float)sinp;
    *cs = (float)cosp;
  } else {
    *sn = (float)cosp;
    // cosf with odd n evaluates the sine polynomial; the table-1 switch only negates cosine coefficients,
    // the sign of the sine polynomial is carried by xs
    *cs = (float)sinp;
  }
}

struct Coef {
  float b0, b1, b2, a1, a2;
};

// UpdateCoefficients, BiQuadFilterNode.cs:149-258
static __device__ Coef rbj(int type, float frequency, float q, float gain, int sample_rate) {
  float w0 = 2.f * 3.14159274f * frequency / (float)sample_rate;  // left to right in float32 (:151)
  float sinW0, cosW0;
  sincosf_libm(w0, &sinW0, &cosW0);
  float alpha = sinW0 / (2.f * q);
  float a0, a1, a2, b0, b1, b2;
  switch (type) {
    case 0: b0 = (1.f - cosW0) / 2.f; b1 = 1.f - cosW0; b2 = (1.f - cosW0) / 2.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 1: b0 = (1.f + cosW0) / 2.f; b1 = -(1.f + cosW0); b2 = (1.f + cosW0) / 2.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 2: b0 = alpha; b1 = 0.f; b2 = -alpha; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 3: b0 = 1.f; b1 = -2.f * cosW0; b2 = 1.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 4: b0 = 1.f - alpha; b1 = -2.f * cosW0; b2 = 1.f + alpha; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 5: {
      float A = (float)pow(10.0, (double)(gain / 40.f));  // MathF.Pow -> powf; double pow rounded to float agrees except on rare ties
      b0 = 1.f + alpha * A; b1 = -2.f * cosW0; b2 = 1.f - alpha * A; a0 = 1.f + alpha / A; a1 = -2.f * cosW0; a2 = 1.f - alpha / A; break;
    }
    case 6: {
      float A = (float)pow(10.0, (double)(gain / 40.f));
      float sqrtA = sqrtf(A);
      float beta = sqrtA / q;
      b0 = A * ((A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0);
      b1 = 2.f * A * ((A - 1.f) - (A + 1.f) * cosW0);
      b2 = A * ((A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0);
      a0 = (A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0;
      a1 = -2.f * ((A - 1.f) + (A + 1.f) * cosW0);
      a2 = (A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0;
      break;
    }
    case 7: {
      float A = (float)pow(10.0, (double)(gain / 40.f));
      float sqrtA = sqrtf(A);
      float beta = sqrtA / q;
      b0 = A * ((A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0);
      b1 = -2.f * A * ((A - 1.f) + (A + 1.f) * cosW0);
      b2 = A * ((A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0);
      a0 = (A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0;
      a1 = 2.f * ((A - 1.f) - (A + 1.f) * cosW0);
      a2 = (A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0;
      break;
    }
    default: b0 = 1.f; b1 = 0.f; b2 = 0.f; a0 = 1.f; a1 = 0.f; a2 = 0.f; break;
  }
  Coef c;
  c.b0 = b0 / a0; c.b1 = b1 / a0; c.b2 = b2 / a0; c.a1 = a1 / a0; c.a2 = a2 / a0;  // :253-257
  return c;
}

__device__ __forceinline__ float clamped_freq(const BiquadJob& job, int64_t n, float nyq) {
  float f = job.freq ? job.freq[n] : job.freq_const;
  return f < 1.f ? 1.f : (f > nyq ? nyq : f);  // Math.Clamp(freq, 1, fs/2)  :123
}
__device__ __forceinline__ float clamped_q(const BiquadJob& job, int64_t n) {
  float q = job.q ? job.q[n] : job.q_const;
  return q > 0.001f ? q : 0.001f;  // Math.Max(0.001f, q)  :124
}


}  // namespace gac
