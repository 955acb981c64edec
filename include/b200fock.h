/*
 * b200fock.h -- C ABI of libb200fock.so: hand-written sm_100a CUDA kernels for a
 * Fock-basis simulator (complex128 state, D^n amplitudes pure / D^2n mixed).
 *
 * The reference (XanaduAI/strawberryfields) has no FFI: its fock backend is NumPy +
 * numba.  Each entry point below replaces one reference compute site; the
 * comment on each cites the reference file:line (paths relative to
 * strawberryfields/backends/).  INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - b200_c128 is an interleaved (re, im) pair of doubles, same memory layout as
 *    numpy complex128 / torch.complex128.
 *  - Every pointer named *_dev is DEVICE memory owned by the caller; the library
 *    never allocates, frees or copies state memory on its own.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Every call is asynchronous on that stream.
 *  - All functions return 0 on success, a positive cudaError_t on a CUDA error,
 *    or a negative B200_E* code on bad arguments; b200_last_error() then holds a
 *    message.  Nothing here falls back to the CPU.
 *  - State tensors are C-order.  A gate on one axis views the state as
 *    [outer, D, inner]; `stride` arguments are element strides of the gate axes.
 *  - Batched gates: `nbatch` leading copies of the state, each `state_batch_stride`
 *    elements apart, using gate table `b * gate_batch_stride` (0 = shared gate).
 */
#ifndef B200FOCK_H
#define B200FOCK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } b200_c128;

#define B200_EINVAL (-1)       /* bad argument */
#define B200_EUNSUPPORTED (-2) /* cutoff / rank outside the compiled range */

#define B200_MAX_CUTOFF 64        /* largest cutoff any kernel accepts (one-mode, diagonal, reductions) */
#define B200_MAX_PAIR_CUTOFF 27   /* two-axis operators: the packed table (b200_packed_size) must fit the
                                     227 KB of shared memory the gate kernels stage it in; b200_gen_gate2
                                     additionally generates BS / MZ / S2 tables up to cutoff 32 only */
#define B200_MAX_FAST_CUTOFF 16   /* cutoffs <= this use the register-blocked kernels */
#define B200_MAX_AXES 24          /* rank limit of the strided gather/reduce kernel */

/* gate kinds for b200_gen_gate1 / b200_gen_diag / b200_gen_gate2 */
enum {
  B200_GATE_DISPLACEMENT = 1, /* fockbackend/ops.py:219-235  (thewalrus displacement) */
  B200_GATE_SQUEEZE = 2,      /* fockbackend/ops.py:238-254  (thewalrus squeezing)    */
  B200_DIAG_ROTATION = 10,    /* fockbackend/ops.py:309-314  exp(i theta n)           */
  B200_DIAG_KERR = 11,        /* fockbackend/ops.py:274-281  exp(i kappa n^2)         */
  B200_DIAG_CROSS_KERR = 12,  /* fockbackend/ops.py:284-293  exp(i kappa n1 n2)       */
  B200_GATE_BEAMSPLITTER = 20,/* fockbackend/ops.py:318-329  (thewalrus beamsplitter) */
  B200_GATE_MZ = 21,          /* fockbackend/ops.py:332-343  (thewalrus mzgate)       */
  B200_GATE_S2 = 22,          /* fockbackend/ops.py:257-271  (thewalrus two_mode_squeezing) */
  B200_CHANNEL_LOSS = 30      /* fockbackend/ops.py:471-490  Kraus sum as one superoperator */
};

/* block structure of a two-axis operator (selection rule) */
enum {
  B200_RULE_SINGLE = 0, /* dense D x D on one axis                                        */
  B200_RULE_SUM = 1,    /* <i j|G|k l> ~ delta(i+j, k+l): BSgate, MZgate (circuit.py:307-335) */
  B200_RULE_DIFF = 2    /* <i j|G|k l> ~ delta(i-j, k-l): S2gate (circuit.py:338-365), loss  */
};

/* ---- library ------------------------------------------------------------------ */
int b200_version(void);
const char* b200_last_error(void);
/* number of b200_c128 entries of a block-packed two-axis operator: D^2 + (D-1)D(2D-1)/3 */
int64_t b200_packed_size(int D);
/* kernel launches issued by this process through the library since load / last reset */
int64_t b200_launch_count(void);
void b200_reset_launch_count(void);
/* let kernels launched on the current device dereference pointers that live on `peer_device`
 * (NVLink peer access; the multi-GPU exchange reads other ranks' shards directly) */
int b200_enable_peer_access(int peer_device);

/* ---- gate tables, generated on the device (reference: thewalrus.fock_gradients
 *      called from fockbackend/ops.py:233,252,266,326,340) -----
 * params_dev: optional DEVICE array [2][nbatch] (row 0 = first parameter); when NULL
 * the scalars (p0, p1) are used for every batch element.                            */
int b200_gen_gate1(int kind, int D, int nbatch, double p0, double p1,
                   const double* params_dev, b200_c128* out_dev /*[nbatch][D][D] out,in*/,
                   void* stream);
/* diagonal gates: out [nbatch][D] (rotation, kerr) or [nbatch][D][D] (cross kerr) */
int b200_gen_diag(int kind, int D, int nbatch, double p0, const double* params_dev,
                  b200_c128* out_dev, void* stream);
/* two-mode gates / loss superoperator in block-packed form [nbatch][b200_packed_size(D)] */
int b200_gen_gate2(int kind, int D, int nbatch, double p0, double p1,
                   const double* params_dev, b200_c128* out_dev, void* stream);
/* C[b] = A[b] * B[b]   (D x D, row-major) -- pre-multiplying consecutive gates on one mode */
int b200_compose_gate1(int D, int nbatch, const b200_c128* A_dev, const b200_c128* B_dev,
                       b200_c128* C_dev, void* stream);
/* fold diagonal phases into a gate table so that a diagonal gate costs no pass:
 *   gate1:  U <- diag(post) U diag(pre)
 *   gate2:  G <- (post1 x post2) G (pre1 x pre2)  on a block-packed table
 * any of pre/post may be NULL (= identity). vectors are [nbatch][D].                */
int b200_fold_diag_gate1(int D, int nbatch, b200_c128* U_dev, const b200_c128* pre_dev,
                         const b200_c128* post_dev, void* stream);
int b200_fold_diag_gate2(int rule, int D, int nbatch, b200_c128* G_dev,
                         const b200_c128* pre1_dev, const b200_c128* pre2_dev,
                         const b200_c128* post1_dev, const b200_c128* post2_dev, void* stream);
/* out[i] = a[i] * (conj_b ? conj(b[i]) : b[i])  -- products of diagonal-gate tables */
int b200_mul_tables(int64_t n, const b200_c128* a_dev, const b200_c128* b_dev, int conj_b,
                    b200_c128* out_dev, void* stream);
/* unpack a block-packed operator into the dense [o1][i1][o2][i2] tensor (tests / host users) */
int b200_unpack_gate2(int rule, int D, const b200_c128* packed_dev, b200_c128* dense_dev,
                      void* stream);

/* ---- gate application (in place) ---------------------------------------------------
 * replaces Circuit.apply_gate_BLAS (fockbackend/circuit.py:118-217): psi'[o,a,i] =
 * sum_b U[a,b] psi[o,b,i] on the view [outer, D, inner].  conj != 0 applies U*
 * (the bra side of rho' = U rho U^dagger).                                           */
int b200_apply_gate1(b200_c128* state_dev, int64_t outer, int D, int64_t inner,
                     const b200_c128* U_dev, int conj, int nbatch, int64_t state_batch_stride,
                     int64_t gate_batch_stride, void* stream);
/* replaces Circuit.apply_twomode_gate + the numba kernels (circuit.py:219-365) and, with
 * B200_RULE_DIFF on the (ket, bra) axes of one mode, Circuit.loss/_apply_channel
 * (circuit.py:65-87,617-621).  stride1/stride2 = element strides of the axes carrying the
 * gate's first / second index; total = elements per batch entry.                      */
int b200_apply_gate2(b200_c128* state_dev, int64_t total, int D, int64_t stride1,
                     int64_t stride2, int rule, const b200_c128* packed_dev, int conj,
                     int nbatch, int64_t state_batch_stride, int64_t gate_batch_stride,
                     void* stream);
/* diagonal gates (circuit.py:160-162,173-175,188-190,202-206): state *= tab[digit(axis1)]
 * (stride2 == 0, tab [D]) or tab[digit1][digit2] (tab [D][D]).                        */
int b200_apply_diag(b200_c128* state_dev, int64_t total, int D, int64_t stride1,
                    int64_t stride2, const b200_c128* tab_dev, int conj, int nbatch,
                    int64_t state_batch_stride, int64_t tab_batch_stride, void* stream);

/* all pending diagonal gates in ONE pass: state[e] *= prod_k op_k(tabs[k][digit_k(e)]),
 * digit_k(e) = (e / strides[k]) % D, op_k = conj when conj_flags[k] != 0.
 * strides / conj_flags are HOST arrays of length naxes (<= B200_MAX_AXES);
 * tabs_dev is [nbatch][naxes][D].                                                      */
int b200_apply_diag_multi(b200_c128* state_dev, int64_t total, int D, int naxes,
                          const int64_t* strides, const int* conj_flags,
                          const b200_c128* tabs_dev, int nbatch, int64_t state_batch_stride,
                          int64_t tab_batch_stride, void* stream);

/* ---- tile pass: several gates per HBM round trip ---------------------------------------
 * A tile is the D x D x D sub-tensor spanned by two axes of element strides
 * stride0 > stride1 (> 1) and the innermost axis (stride 1), every other index fixed.
 * The kernel stages tiles in shared memory, applies `nops` operators IN ORDER to tile
 * axes 0 (stride0), 1 (stride1), 2 (innermost), and writes the tiles back with tile axis
 * k placed at tile position out_perm[k] (identity = {0,1,2}); the permutation lets the
 * caller rotate which logical mode is innermost at no extra cost.
 *   kind B200_RULE_SINGLE: dense D x D table on axis1
 *   kind B200_RULE_SUM / B200_RULE_DIFF: block-packed table on (axis1, axis2)
 *   kind B200_TILE_DIAG: table [D] on axis1
 * conj != 0 applies the complex-conjugated table (bra side of a mixed state).
 * coef_dev: [nbatch][coef_count] arena; each op names its own range (coef_offset).
 * Cutoffs 2..B200_MAX_FAST_CUTOFF; fails with B200_EUNSUPPORTED when the tiles plus the
 * operators' tables exceed shared memory (b200_tile_smem_bytes tells in advance).      */
#define B200_TILE_DIAG 3
#define B200_TILE_MAX_OPS 16
typedef struct {
  int kind;
  int axis1, axis2;
  int conj;
  int64_t coef_offset;
} b200_tile_op;
int b200_tile_groups(int D);                          /* tiles staged per CTA */
int64_t b200_tile_smem_bytes(int D, int64_t coef_count);
int b200_apply_tile_pass(b200_c128* state_dev, int64_t total, int D, int64_t stride0,
                         int64_t stride1, const b200_tile_op* ops, int nops,
                         const int* out_perm, const b200_c128* coef_dev, int64_t coef_count,
                         int nbatch, int64_t state_batch_stride, int64_t coef_batch_stride,
                         void* stream);

/* ---- strided gather / product / reduce (state preparation, partial traces, marginals,
 *      reduced density matrices; replaces the einsum helpers fockbackend/ops.py:110-198,
 *      circuit.py:393-473, backend.py:219-259, states.py:580-642)
 *   C[c(o)] = sum_r A[a(o) + ta(r)] * op(B[b(o) + tb(r)])      (B NULL -> factor 1)
 * o runs over the `n_out_axes` output digits (extents out_ext, C-order, last fastest),
 * r over the `n_red_axes` reduction digits.  flags: bit0 conjugate B, bit1 write only the
 * real part to a double array, bit2 A is a double array (B must be NULL).
 * part_dev: scratch of >= n_out * 64 b200_c128 used when the reduction is split.      */
typedef struct {
  int n_out_axes;
  int n_red_axes;
  int32_t out_ext[B200_MAX_AXES];
  int64_t out_sa[B200_MAX_AXES], out_sb[B200_MAX_AXES], out_sc[B200_MAX_AXES];
  int32_t red_ext[B200_MAX_AXES];
  int64_t red_ta[B200_MAX_AXES], red_tb[B200_MAX_AXES];
  int64_t base_a, base_b, base_c;
} b200_gather_desc;
int b200_gather_reduce(const b200_gather_desc* desc, const void* A_dev, const b200_c128* B_dev,
                       void* C_dev, int flags, b200_c128* part_dev, void* stream);

/* ---- elementwise helpers --------------------------------------------------------------- */
/* out[b][j][i] = f[b][j] * in[b][i], 0 <= j < nf, 0 <= i < n_in: a product factor (|0> under its pending
 * single-mode operator: nf = D; a rank-one (ket, bra) factor of a density matrix: nf = D * D) becomes the new
 * outermost axis of the tensor -- the lazy vacuum's replacement for np.tensordot / the reference's full-size
 * allocation at begin_circuit (circuit.py:89-116).  f_batch_stride = 0: one factor for every batch entry.     */
int b200_outer_axis(const b200_c128* in_dev, const b200_c128* f_dev, b200_c128* out_dev, int64_t n_in, int nf,
                    int nbatch, int64_t in_batch_stride, int64_t f_batch_stride, void* stream);
int b200_fill_zero(b200_c128* dev, int64_t n, void* stream);
int b200_set_element(b200_c128* dev, int64_t index, double re, double im, void* stream);
/* probs[i] = |psi[i]|^2  -- states.py:596-598 */
int b200_abs2(const b200_c128* psi_dev, double* probs_dev, int64_t n, void* stream);
/* out (double, DEVICE) = sum |psi|^2 ; two-stage, deterministic.  part_dev: >= 4096 doubles.
 * circuit.py:367-371 */
int b200_norm2(const b200_c128* psi_dev, int64_t n, double* out_dev, double* part_dev,
               void* stream);
/* state *= (re, im) / (*divisor_dev if divisor_dev else 1)   [sqrt_div: divide by sqrt] */
int b200_scale(b200_c128* dev, int64_t n, double re, double im, const double* divisor_dev,
               int sqrt_div, void* stream);

/* ---- single-mode reduced density matrix / photon-number marginal of a ket in ONE read of the state
 *      (reference: backend.py:219-253, states.py:613-642, circuit.py:675-677 -- einsum / tensordot on the host).
 * The state is viewed as [outer, D, inner] along the kept mode's axis.  diag_only = 0: out_dev receives the
 * Hermitian D x D matrix rho[a][b] = sum_r psi[.., a, ..] conj(psi[.., b, ..]) (b200_c128, [nbatch][D][D]);
 * diag_only = 1: the marginal probabilities (double, [nbatch][D]).  part_dev: scratch of
 * b200_gram1_part_doubles(D, nbatch) doubles.  Deterministic (fixed reduction order).  Cutoffs 1..12 (matrix),
 * 1..16 (marginal).                                                                                          */
int64_t b200_gram1_part_doubles(int D, int nbatch);
int b200_gram1(const b200_c128* psi_dev, int64_t outer, int D, int64_t inner, int diag_only, void* out_dev,
               double* part_dev, int nbatch, int64_t state_batch_stride, void* stream);

/* ---- multi-GPU axis exchange (the reference has no distributed path) ------------
 * One launch performs a rank's whole share of the all-to-all that swaps the sharded leading
 * axes of the state with local ones.  For every source s the copy is
 *   dst[s][dst_base[s] + sum_j i_j*ds[j] + r] = src[s][src_base[s] + sum_j i_j*ss[j] + r],
 * 0 <= i_j < ext[j], 0 <= r < run (elements, contiguous on both sides).  src[] / dst[] may be
 * peer-device pointers (CUDA IPC + peer access): "pull" passes the peers' old shards as src,
 * "push" the peers' new shards as dst.  The payload moves through shared memory with the bulk
 * copy engine (cp.async.bulk global->shared->global), sources interleaved starting at
 * `first_src`.  local_src >= 0 (then == first_src): that source and its destination are both
 * local; its block is copied with plain loads / stores by the CTAs the link does not need
 * (-1: every source goes through the bulk path).  bulk_ctas: CTAs driving the bulk path,
 * 0 = default (32 next to a local block, else one per SM).                                       */
#define B200_XCHG_MAX_AXES 12
#define B200_XCHG_MAX_PEERS 32
typedef struct {
  int n_axes;
  int n_src;
  int first_src;
  int32_t ext[B200_XCHG_MAX_AXES];
  int64_t ss[B200_XCHG_MAX_AXES], ds[B200_XCHG_MAX_AXES];
  int64_t run;
  const void* src[B200_XCHG_MAX_PEERS];
  void* dst[B200_XCHG_MAX_PEERS];
  int64_t src_base[B200_XCHG_MAX_PEERS], dst_base[B200_XCHG_MAX_PEERS];
} b200_xchg_desc;
int b200_exchange_copy(const b200_xchg_desc* desc, int local_src, int bulk_ctas, void* stream);

/* Device-side barrier between the ranks of a sharded state, stream-ordered (no host
 * synchronisation): flags[r] points at rank r's array of n_ranks uint64 counters (peer-mapped,
 * zero-initialised); every rank calls with the same, strictly increasing `epoch`.  A rank that
 * waits longer than timeout_s traps (the launch fails) instead of hanging the GPU.             */
typedef struct {
  int n_ranks;
  int rank;
  void* flags[B200_XCHG_MAX_PEERS];
} b200_peer_flags;
int b200_peer_barrier(const b200_peer_flags* pf, uint64_t epoch, double timeout_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FOCK_H */
