// Gate-table generation on the device (K9 in SURVEY.md section 2).
//
// The reference obtains these tensors from thewalrus.fock_gradients
// (strawberryfields/backends/fockbackend/ops.py:233,252,266,326,340; thewalrus 0.22.0,
// source not in the reference tree).  The recursions below are the published ones
//  re-derived for a block-packed layout: a two-mode tensor
// Z[m,n,p,q] = <m n|G|p q> is non-zero only for m+n = p+q (beamsplitter, MZ) or
// m-n = p-q (two-mode squeezing), so it is stored as W[m][n][p] with q implied, directly
// in the packed form the apply kernel consumes -- no D^4 work array is ever built.
#include "common.cuh"

namespace b200 {

// ---- single-mode dense gates ----------------------------------------------------------
// one thread per batch element; D^2 complex entries, strictly sequential recursion.
__global__ void k_gen_gate1(int kind, int D, int nbatch, double p0, double p1,
                            const double* __restrict__ params, cplx* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbatch) return;
  if (params) {
    p0 = params[b];
    p1 = params[nbatch + b];
  }
  cplx* T = out + (size_t)b * D * D;
  double sn, cs;
  sincos(p1, &sn, &cs);
  if (kind == B200_GATE_DISPLACEMENT) {
    // D[0,0] = exp(-r^2/2); D[m,0] = alpha/sqrt(m) D[m-1,0];
    // D[m,n] = -conj(alpha)/sqrt(n) D[m,n-1] + sqrt(m/n) D[m-1,n-1]
    cplx alpha = make_double2(p0 * cs, p0 * sn);
    cplx nac = make_double2(-alpha.x, alpha.y);
    T[0] = make_double2(exp(-0.5 * p0 * p0), 0.0);
    for (int m = 1; m < D; ++m) T[m * D] = cscale(cmul(alpha, T[(m - 1) * D]), 1.0 / sqrt((double)m));
    for (int m = 0; m < D; ++m)
      for (int n = 1; n < D; ++n) {
        double rn = 1.0 / sqrt((double)n);
        cplx v = cscale(cmul(nac, T[m * D + n - 1]), rn);
        if (m > 0) v = cadd(v, cscale(T[(m - 1) * D + n - 1], sqrt((double)m) * rn));
        T[m * D + n] = v;
      }
  } else {  // B200_GATE_SQUEEZE
    // t = e^{i theta} tanh r, s = sech r; S[0,0] = sqrt(s); S[m,0] = -t sqrt((m-1)/m) S[m-2,0];
    // S[m,n] = conj(t) sqrt((n-1)/n) S[m,n-2] + s sqrt(m/n) S[m-1,n-1]   (m+n even)
    double th = tanh(p0), sech = 1.0 / cosh(p0);
    cplx t = make_double2(th * cs, th * sn);
    cplx nt = make_double2(-t.x, -t.y), ct = cconj(t);
    for (int i = 0; i < D * D; ++i) T[i] = make_double2(0.0, 0.0);
    T[0] = make_double2(sqrt(sech), 0.0);
    for (int m = 2; m < D; m += 2)
      T[m * D] = cscale(cmul(nt, T[(m - 2) * D]), sqrt((double)(m - 1)) / sqrt((double)m));
    for (int m = 0; m < D; ++m)
      for (int n = 1; n < D; ++n) {
        if ((m + n) & 1) continue;
        double rn = 1.0 / sqrt((double)n);
        cplx v = make_double2(0.0, 0.0);
        if (n >= 2) v = cscale(cmul(ct, T[m * D + n - 2]), sqrt((double)(n - 1)) * rn);
        if (m > 0) v = cadd(v, cscale(T[(m - 1) * D + n - 1], sech * sqrt((double)m) * rn));
        T[m * D + n] = v;
      }
  }
}

// ---- diagonal gates ----------------------------------------------------------------------
__global__ void k_gen_diag(int kind, int D, int nbatch, double p0, const double* __restrict__ params,
                           cplx* __restrict__ out) {
  int per = (kind == B200_DIAG_CROSS_KERR) ? D * D : D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)per * nbatch) return;
  int b = (int)(i / per), e = (int)(i % per);
  if (params) p0 = params[b];
  double k;
  if (kind == B200_DIAG_ROTATION) k = (double)e;
  else if (kind == B200_DIAG_KERR) k = (double)e * (double)e;
  else k = (double)(e / D) * (double)(e % D);
  double sn, cs;
  sincos(p0 * k, &sn, &cs);
  out[i] = make_double2(cs, sn);
}

// ---- two-mode gates, block-packed ---------------------------------------------------------
struct PackedView {
  cplx* W;
  int D;
  // SUM layout: block = m + n ; DIFF layout: block = m - n + D - 1.  Entry (m, n, p).
  __device__ __forceinline__ int idx_sum(int m, int n, int p) const {
    if (m < 0 || n < 0 || p < 0) return -1;
    int b = m + n, lo = blk_lo(b, D), c = blk_size(b, D);
    if (p < lo || p >= lo + c) return -1;
    return blk_off(b, D) + (m - lo) * c + (p - lo);
  }
  __device__ __forceinline__ int idx_diff(int m, int n, int p) const {
    if (m < 0 || n < 0 || p < 0) return -1;
    int b = m - n + D - 1, lo = blk_lo(b, D), c = blk_size(b, D);
    if (p < lo || p >= lo + c) return -1;
    return blk_off(b, D) + (m - lo) * c + (p - lo);
  }
  __device__ __forceinline__ cplx get_sum(int m, int n, int p) const {
    int i = idx_sum(m, n, p);
    return i < 0 ? make_double2(0.0, 0.0) : W[i];
  }
  __device__ __forceinline__ cplx get_diff(int m, int n, int p) const {
    int i = idx_diff(m, n, p);
    return i < 0 ? make_double2(0.0, 0.0) : W[i];
  }
};

// Passive two-mode unitary with 2x2 mode transformation [[u00, u01], [u10, u11]]:
//   q = 0:  W[m,n,p=m+n] = u00 sqrt(m/p) W[m-1,n,p-1] + u10 sqrt(n/p) W[m,n-1,p-1]
//   q > 0:  W[m,n,p]     = u01 sqrt(m/q) W[m-1,n,p]   + u11 sqrt(n/q) W[m,n-1,p],  q = m+n-p
// One CTA of D x D threads (p = threadIdx.y, m = threadIdx.x) per batch element; the
// wavefront t = m + n advances with one __syncthreads per step.
__global__ void k_gen_passive(int kind, int D, int nbatch, double p0, double p1,
                              const double* __restrict__ params, cplx* __restrict__ out) {
  int b = blockIdx.x;
  if (params) {
    p0 = params[b];
    p1 = params[nbatch + b];
  }
  cplx u00, u01, u10, u11;
  if (kind == B200_GATE_BEAMSPLITTER) {
    double st, ct, sp, cp;
    sincos(p0, &st, &ct);
    sincos(p1, &sp, &cp);
    cplx s = make_double2(st * cp, st * sp);
    u00 = make_double2(ct, 0.0);
    u01 = make_double2(-s.x, s.y);  // -conj(s)
    u10 = s;
    u11 = make_double2(ct, 0.0);
  } else {  // MZ: v = e^{i phi_in}, u = e^{i phi_ex}
    double sv, cv, su, cu;
    sincos(p0, &sv, &cv);
    sincos(p1, &su, &cu);
    cplx v = make_double2(cv, sv), u = make_double2(cu, su);
    cplx vm1 = make_double2(v.x - 1.0, v.y), vp1 = make_double2(v.x + 1.0, v.y);
    cplx ivp1 = make_double2(-vp1.y, vp1.x);  // i (v + 1)
    u00 = cscale(cmul(vm1, u), 0.5);           // (v-1) u / 2
    u01 = cscale(ivp1, 0.5);                   // i (v+1) / 2
    u10 = cscale(cmul(ivp1, u), 0.5);          // i (v+1) u / 2
    u11 = make_double2(0.5 * (1.0 - v.x), -0.5 * v.y);  // (1-v)/2
  }
  PackedView V{out + (size_t)b * packed_size(D), D};
  int m = threadIdx.x, p = threadIdx.y;
  int tid = p * D + m, nthr = D * D;
  for (int i = tid; i < packed_size(D); i += nthr) V.W[i] = make_double2(0.0, 0.0);
  __syncthreads();
  if (tid == 0) V.W[0] = make_double2(1.0, 0.0);
  __syncthreads();
  // q = 0 line: p = t = m + n
  for (int t = 1; t < D; ++t) {
    if (p == 0 && m <= t) {
      int n = t - m;
      double rt = 1.0 / sqrt((double)t);
      cplx a = cscale(cmul(u00, V.get_sum(m - 1, n, t - 1)), sqrt((double)m) * rt);
      cplx c = cscale(cmul(u10, V.get_sum(m, n - 1, t - 1)), sqrt((double)n) * rt);
      V.W[V.idx_sum(m, n, t)] = cadd(a, c);
    }
    __syncthreads();
  }
  // q > 0: fixed p per thread row, wavefront over t = m + n
  for (int t = 1; t <= 2 * D - 2; ++t) {
    int n = t - m, q = t - p;
    if (n >= 0 && n < D && q >= 1 && q < D) {
      double rq = 1.0 / sqrt((double)q);
      cplx a = cscale(cmul(u01, V.get_sum(m - 1, n, p)), sqrt((double)m) * rq);
      cplx c = cscale(cmul(u11, V.get_sum(m, n - 1, p)), sqrt((double)n) * rq);
      V.W[V.idx_sum(m, n, p)] = cadd(a, c);
    }
    __syncthreads();
  }
}

// Two-mode squeezing, s = sech r, e = e^{i theta} tanh r:
//   W[0,0,0] = s; W[n,n,0] = e W[n-1,n-1,0];
//   q = 0, p = m-n > 0: W[m,n,p] = s sqrt(m/p) W[m-1,n,p-1]
//   q > 0 (q = p-(m-n)): W[m,n,p] = s sqrt(n/q) W[m,n-1,p] - conj(e) sqrt(p/q) W[m,n,p-1]
// One CTA of D x D threads (m = threadIdx.x, n = threadIdx.y) per batch element.
__global__ void k_gen_s2(int D, int nbatch, double p0, double p1, const double* __restrict__ params,
                         cplx* __restrict__ out) {
  int b = blockIdx.x;
  if (params) {
    p0 = params[b];
    p1 = params[nbatch + b];
  }
  double sn, cs;
  sincos(p1, &sn, &cs);
  double th = tanh(p0), s = 1.0 / cosh(p0);
  cplx e = make_double2(th * cs, th * sn);
  cplx nce = make_double2(-e.x, e.y);  // -conj(e)
  PackedView V{out + (size_t)b * packed_size(D), D};
  int m = threadIdx.x, n = threadIdx.y;
  int tid = n * D + m, nthr = D * D;
  for (int i = tid; i < packed_size(D); i += nthr) V.W[i] = make_double2(0.0, 0.0);
  __syncthreads();
  if (tid == 0) {
    V.W[V.idx_diff(0, 0, 0)] = make_double2(s, 0.0);
    for (int k = 1; k < D; ++k) V.W[V.idx_diff(k, k, 0)] = cmul(e, V.W[V.idx_diff(k - 1, k - 1, 0)]);
  }
  __syncthreads();
  // q = 0: chains in m for fixed n (thread row m == 0 walks the chain of its n)
  if (m == 0) {
    for (int mm = n + 1; mm < D; ++mm) {
      int p = mm - n;
      V.W[V.idx_diff(mm, n, p)] =
          cscale(V.get_diff(mm - 1, n, p - 1), s * sqrt((double)mm) / sqrt((double)p));
    }
  }
  __syncthreads();
  for (int q = 1; q < D; ++q) {
    int p = q + m - n;
    if (p >= 0 && p < D) {
      double rq = 1.0 / sqrt((double)q);
      cplx a = cscale(V.get_diff(m, n - 1, p), s * sqrt((double)n) * rq);
      cplx c = cscale(cmul(nce, V.get_diff(m, n, p - 1)), sqrt((double)p) * rq);
      V.W[V.idx_diff(m, n, p)] = cadd(a, c);
    }
    __syncthreads();
  }
}

// Loss channel as ONE superoperator on the (ket, bra) axes of a mode:
//   rho'[a,d] = sum_l (1-T)^l T^{(a+d)/2} sqrt(C(a+l,l) C(d+l,l)) rho[a+l, d+l]
// (= sum_l E_l rho E_l^dagger with the Kraus operators of fockbackend/ops.py:471-490).
// DIFF block (a - d fixed): row member a = lo + i, column member a + l = lo + j, l = j - i >= 0.
__global__ void k_gen_loss(int D, int nbatch, double T0, const double* __restrict__ params,
                           cplx* __restrict__ out) {
  int P = packed_size(D);
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)P * nbatch) return;
  int bt = (int)(g / P), e = (int)(g % P);
  double T = params ? params[bt] : T0;
  // locate block containing packed index e
  int b = 0;
  while (b < 2 * D - 2 && blk_off(b + 1, D) <= e) ++b;
  int c = blk_size(b, D), lo = blk_lo(b, D), dd = b - (D - 1);
  int r = e - blk_off(b, D);
  int i = r / c, j = r % c;
  double val = 0.0;
  if (j >= i) {
    int l = j - i, a = lo + i, d = lo + i - dd;
    // sqrt(C(a+l,l) C(d+l,l)) by running products (exact in double for these sizes)
    double ca = 1.0, cd = 1.0;
    for (int k = 1; k <= l; ++k) {
      ca = ca * (double)(a + k) / (double)k;
      cd = cd * (double)(d + k) / (double)k;
    }
    val = pow(1.0 - T, (double)l) * pow(T, 0.5 * (double)(a + d)) * sqrt(ca * cd);
  }
  out[g] = make_double2(val, 0.0);
}

// ---- composition helpers -----------------------------------------------------------------
__global__ void k_compose_gate1(int D, int nbatch, const cplx* __restrict__ A, const cplx* __restrict__ B,
                                cplx* __restrict__ C) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)D * D * nbatch) return;
  int b = (int)(g / (D * D)), e = (int)(g % (D * D));
  int i = e / D, j = e % D;
  const cplx* a = A + (size_t)b * D * D;
  const cplx* bb = B + (size_t)b * D * D;
  cplx acc = make_double2(0.0, 0.0);
  for (int k = 0; k < D; ++k) cfma(acc, a[i * D + k], bb[k * D + j]);
  C[g] = acc;
}

__global__ void k_fold_diag_gate1(int D, int nbatch, cplx* __restrict__ U, const cplx* __restrict__ pre,
                                  const cplx* __restrict__ post) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)D * D * nbatch) return;
  int b = (int)(g / (D * D)), e = (int)(g % (D * D));
  int i = e / D, j = e % D;
  cplx v = U[g];
  if (pre) v = cmul(v, pre[(size_t)b * D + j]);
  if (post) v = cmul(post[(size_t)b * D + i], v);
  U[g] = v;
}

__device__ __forceinline__ void member_kl(int rule, int D, int b, int lo, int mem, int& k, int& l) {
  k = lo + mem;
  l = (rule == B200_RULE_SUM) ? (b - lo - mem) : (lo + mem - (b - (D - 1)));
}

__global__ void k_fold_diag_gate2(int rule, int D, int nbatch, cplx* __restrict__ G,
                                  const cplx* __restrict__ pre1, const cplx* __restrict__ pre2,
                                  const cplx* __restrict__ post1, const cplx* __restrict__ post2) {
  int P = packed_size(D);
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)P * nbatch) return;
  int bt = (int)(g / P), e = (int)(g % P);
  int b = 0;
  while (b < 2 * D - 2 && blk_off(b + 1, D) <= e) ++b;
  int c = blk_size(b, D), lo = blk_lo(b, D);
  int r = e - blk_off(b, D);
  int ko, lo_, ki, li;
  member_kl(rule, D, b, lo, r / c, ko, lo_);
  member_kl(rule, D, b, lo, r % c, ki, li);
  cplx v = G[g];
  size_t o = (size_t)bt * D;
  if (pre1) v = cmul(v, pre1[o + ki]);
  if (pre2) v = cmul(v, pre2[o + li]);
  if (post1) v = cmul(post1[o + ko], v);
  if (post2) v = cmul(post2[o + lo_], v);
  G[g] = v;
}

__global__ void k_unpack_gate2(int rule, int D, const cplx* __restrict__ packed, cplx* __restrict__ dense) {
  // dense[o1][i1][o2][i2]
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int D2 = D * D;
  if (g >= D2 * D2) return;
  int o1 = g / (D2 * D), i1 = (g / D2) % D, o2 = (g / D) % D, i2 = g % D;
  int bo = (rule == B200_RULE_SUM) ? (o1 + o2) : (o1 - o2 + D - 1);
  int bi = (rule == B200_RULE_SUM) ? (i1 + i2) : (i1 - i2 + D - 1);
  cplx v = make_double2(0.0, 0.0);
  if (bo == bi) {
    int lo = blk_lo(bo, D), c = blk_size(bo, D);
    v = packed[blk_off(bo, D) + (o1 - lo) * c + (i1 - lo)];
  }
  dense[g] = v;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_gen_gate1(int kind, int D, int nbatch, double p0, double p1, const double* params_dev,
                   b200_c128* out_dev, void* stream) {
  B200_CHECK_ARG(kind == B200_GATE_DISPLACEMENT || kind == B200_GATE_SQUEEZE, "gen_gate1: bad kind");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF && nbatch >= 1 && out_dev, "gen_gate1: bad size");
  int thr = 32;
  k_gen_gate1<<<(nbatch + thr - 1) / thr, thr, 0, (cudaStream_t)stream>>>(kind, D, nbatch, p0, p1, params_dev,
                                                                           (cplx*)out_dev);
  return cuda_status("gen_gate1");
}

int b200_gen_diag(int kind, int D, int nbatch, double p0, const double* params_dev, b200_c128* out_dev,
                  void* stream) {
  B200_CHECK_ARG(kind == B200_DIAG_ROTATION || kind == B200_DIAG_KERR || kind == B200_DIAG_CROSS_KERR,
                 "gen_diag: bad kind");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF && nbatch >= 1 && out_dev, "gen_diag: bad size");
  long long n = (long long)(kind == B200_DIAG_CROSS_KERR ? D * D : D) * nbatch;
  int thr = 128;
  k_gen_diag<<<(unsigned)((n + thr - 1) / thr), thr, 0, (cudaStream_t)stream>>>(kind, D, nbatch, p0, params_dev,
                                                                                 (cplx*)out_dev);
  return cuda_status("gen_diag");
}

int b200_gen_gate2(int kind, int D, int nbatch, double p0, double p1, const double* params_dev,
                   b200_c128* out_dev, void* stream) {
  B200_CHECK_ARG(D >= 1 && nbatch >= 1 && out_dev, "gen_gate2: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == B200_CHANNEL_LOSS) {
    B200_CHECK_ARG(D <= B200_MAX_CUTOFF, "gen_gate2: cutoff too large");
    long long n = (long long)packed_size(D) * nbatch;
    int thr = 128;
    k_gen_loss<<<(unsigned)((n + thr - 1) / thr), thr, 0, st>>>(D, nbatch, p0, params_dev, (cplx*)out_dev);
    return cuda_status("gen_loss");
  }
  if (D > 32) return fail(B200_EUNSUPPORTED, "%s", "gen_gate2: two-mode gate tables support cutoff <= 32");
  dim3 blk(D, D);
  if (kind == B200_GATE_BEAMSPLITTER || kind == B200_GATE_MZ)
    k_gen_passive<<<nbatch, blk, 0, st>>>(kind, D, nbatch, p0, p1, params_dev, (cplx*)out_dev);
  else if (kind == B200_GATE_S2)
    k_gen_s2<<<nbatch, blk, 0, st>>>(D, nbatch, p0, p1, params_dev, (cplx*)out_dev);
  else
    return fail(B200_EINVAL, "%s", "gen_gate2: bad kind");
  return cuda_status("gen_gate2");
}

int b200_compose_gate1(int D, int nbatch, const b200_c128* A_dev, const b200_c128* B_dev, b200_c128* C_dev,
                       void* stream) {
  B200_CHECK_ARG(D >= 1 && nbatch >= 1 && A_dev && B_dev && C_dev, "compose_gate1: bad args");
  B200_CHECK_ARG(C_dev != A_dev && C_dev != B_dev, "compose_gate1: output must not alias an input");
  long long n = (long long)D * D * nbatch;
  int thr = 128;
  k_compose_gate1<<<(unsigned)((n + thr - 1) / thr), thr, 0, (cudaStream_t)stream>>>(
      D, nbatch, (const cplx*)A_dev, (const cplx*)B_dev, (cplx*)C_dev);
  return cuda_status("compose_gate1");
}

int b200_fold_diag_gate1(int D, int nbatch, b200_c128* U_dev, const b200_c128* pre_dev,
                         const b200_c128* post_dev, void* stream) {
  B200_CHECK_ARG(D >= 1 && nbatch >= 1 && U_dev, "fold_diag_gate1: bad args");
  long long n = (long long)D * D * nbatch;
  int thr = 128;
  k_fold_diag_gate1<<<(unsigned)((n + thr - 1) / thr), thr, 0, (cudaStream_t)stream>>>(
      D, nbatch, (cplx*)U_dev, (const cplx*)pre_dev, (const cplx*)post_dev);
  return cuda_status("fold_diag_gate1");
}

int b200_fold_diag_gate2(int rule, int D, int nbatch, b200_c128* G_dev, const b200_c128* pre1_dev,
                         const b200_c128* pre2_dev, const b200_c128* post1_dev, const b200_c128* post2_dev,
                         void* stream) {
  B200_CHECK_ARG(rule == B200_RULE_SUM || rule == B200_RULE_DIFF, "fold_diag_gate2: bad rule");
  B200_CHECK_ARG(D >= 1 && nbatch >= 1 && G_dev, "fold_diag_gate2: bad args");
  long long n = (long long)packed_size(D) * nbatch;
  int thr = 128;
  k_fold_diag_gate2<<<(unsigned)((n + thr - 1) / thr), thr, 0, (cudaStream_t)stream>>>(
      rule, D, nbatch, (cplx*)G_dev, (const cplx*)pre1_dev, (const cplx*)pre2_dev, (const cplx*)post1_dev,
      (const cplx*)post2_dev);
  return cuda_status("fold_diag_gate2");
}

int b200_unpack_gate2(int rule, int D, const b200_c128* packed_dev, b200_c128* dense_dev, void* stream) {
  B200_CHECK_ARG(rule == B200_RULE_SUM || rule == B200_RULE_DIFF, "unpack_gate2: bad rule");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF && packed_dev && dense_dev, "unpack_gate2: bad args");
  int n = D * D * D * D, thr = 128;
  k_unpack_gate2<<<(n + thr - 1) / thr, thr, 0, (cudaStream_t)stream>>>(rule, D, (const cplx*)packed_dev,
                                                                         (cplx*)dense_dev);
  return cuda_status("unpack_gate2");
}

}  // extern "C"
