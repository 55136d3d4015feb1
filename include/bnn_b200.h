/*
 * bnn_b200.h -- C ABI of the B200-native hot path of Alaya-in-Matrix/BNN.
 *
 * The reference has no FFI: its "operator API" is the Python class BNN
 * (BNN.py:6-37) and the bodies of sample / sample_predict / predict_mv /
 * validate in its five subclasses.  Each entry point below replaces one of
 * those bodies; the citation after "replaces:" is the reference code a
 * maintainer would swap for a ctypes call.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the name ends in _host;
 *     the caller owns all memory; nothing is allocated behind the caller's
 *     back except inside the *_host convenience calls (one workspace per
 *     device, cached for the process and freed by bnn_release_workspace());
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     device entry points are asynchronous and never synchronise;
 *   - return value: 0 = ok, <0 = error; bnn_last_error() gives the message;
 *     the library never aborts the process and has NO CPU fallback;
 *   - all floating-point data are IEEE binary32 unless typed double.
 *
 * Packed parameter layout ("bank"): one network = one flat vector of
 * bnn_param_stride() floats.  Linear layer l (in_l -> out_l) occupies
 * out_l rows of ld_l = roundup4(in_l + 1) floats, row-major, starting at
 * bnn_layer_offset(l): columns [0, in_l) are the weight row, column in_l is
 * the bias (bias-last, as BNN_BBB.py:19,30-31,46-47 stores mu/rho), the
 * remaining <=3 columns are zero padding that keeps every row 16-byte aligned
 * for the bulk-copy (TMA) engine.  A bank of S networks is [S, stride].
 */
#ifndef BNN_B200_H_
#define BNN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNN_ABI_VERSION 2
#define BNN_MAX_LINEAR 8       /* Linear layers per network (hidden + output) */
#define BNN_MAX_WIDTH 1024     /* widest layer / input the forward kernels take */
#define BNN_MAX_FUSED_NOUT 64  /* nout limit of the fused forward+moments kernel */

enum { BNN_ACT_RELU = 0, BNN_ACT_TANH = 1, BNN_ACT_SIGMOID = 2, BNN_ACT_IDENTITY = 3 };

/* util.py:8-33 `NN(dim, act, num_hiddens, nout)`: dims = {dim, h1, .., hL, nout} */
typedef struct BnnMlpDesc {
  int32_t n_linear;                  /* L + 1 */
  int32_t dims[BNN_MAX_LINEAR + 1];
  int32_t act;                       /* BNN_ACT_* applied after every Linear but the last */
} BnnMlpDesc;

/* How the per-sample, per-input-unit scaling of BNN_Dropout.py:70-75 /
 * BNN_CDropout.py:59-65 reaches the forward kernel.  The reference scales
 * weight COLUMNS of a deep copy; scaling the layer's INPUT units is the same
 * product and lets all S samples share one set of base weights.            */
enum {
  BNN_MASK_NONE = 0,
  BNN_MASK_EXTERNAL = 1,   /* mask tensor [S, mask_len] supplied (reference-parity mode) */
  BNN_MASK_BERNOULLI = 2,  /* generated in-kernel: Philox, keep iff u < 1-p          */
  BNN_MASK_CONCRETE = 3    /* generated in-kernel: relaxed Bernoulli, T = temperature */
};

typedef struct BnnMaskSpec {
  int32_t mode;                         /* BNN_MASK_* */
  int32_t layer_masked[BNN_MAX_LINEAR]; /* 1 if Linear l's inputs are scaled */
  const float* mask;                    /* EXTERNAL: [S, mask_len], layers concatenated in order */
  int64_t mask_stride;                  /* floats between samples (>= mask_len) */
  float p_drop[BNN_MAX_LINEAR];         /* BERNOULLI/CONCRETE: drop probability per layer */
  float temperature;                    /* CONCRETE (reference: 0.1, BNN_CDropout.py:64) */
  uint64_t seed;                        /* Philox key */
  int64_t sample_offset;                /* global index of sample 0 (multi-GPU shards) */
} BnnMaskSpec;

/* Precision of the dense layers of the forward kernels. */
enum {
  BNN_PREC_FP32 = 0,    /* FP32 FFMA on the SIMT pipe: parity grade (1e-5) */
  BNN_PREC_TF32X3 = 1,  /* tcgen05 kind::tf32, 3-pass error-compensated (~FP32) */
  BNN_PREC_TF32 = 2,    /* tcgen05 kind::tf32 single pass (looser bound) */
  BNN_PREC_TF32_BF16 = 3 /* tcgen05: TF32 main term + two BF16 (kind::f16) correction terms, ~FP32 (2/3 of the tensor work of TF32X3) */
};

typedef struct BnnForwardArgs {
  BnnMlpDesc desc;
  const float* X;        /* [N, dims[0]] row-major */
  int64_t N;
  const float* W;        /* bank [S, w_stride] (w_stride = 0: one shared network) */
  int64_t w_stride;
  int64_t S;
  BnnMaskSpec mask;
  int32_t precision;     /* BNN_PREC_* */
  /* outputs -- any may be NULL */
  float* pred;           /* [S, N, nout]  (sample_predict) */
  float* mean;           /* [N, nout]     (predict_mv)     */
  float* var;            /* [N, nout]  unbiased, S-1       */
  double* part_mean;     /* [N, nout]  this shard's mean   (multi-GPU merge input) */
  double* part_m2;       /* [N, nout]  this shard's sum of squared deviations      */
  void* workspace;       /* >= bnn_forward_workspace_bytes() bytes, 16-byte aligned */
  int64_t workspace_bytes;
} BnnForwardArgs;

/* ---- library ------------------------------------------------------------ */
int32_t bnn_abi_version(void);
const char* bnn_last_error(void);
/* number of SMs of the current device (grid sizing is a multiple of this) */
int32_t bnn_sm_count(void);
/* bind the calling thread to a device (the library carries its own static CUDA
 * runtime; call this with the device your buffers live on before other calls) */
int32_t bnn_set_device(int32_t device);

/* ---- layout (host arithmetic, no GPU) ----------------------------------- */
int64_t bnn_param_stride(const BnnMlpDesc* d);                 /* floats per network */
int64_t bnn_layer_offset(const BnnMlpDesc* d, int32_t l);      /* first float of Linear l */
int32_t bnn_layer_ld(const BnnMlpDesc* d, int32_t l);          /* row pitch of Linear l   */
int64_t bnn_mask_len(const BnnMlpDesc* d, const int32_t* layer_masked);
int64_t bnn_forward_workspace_bytes(const BnnForwardArgs* a);

/* ---- (a)+(e) forward and predictive moments ----------------------------- */
/* replaces: the `for i in range(S): pred[i] = nns[i](X)` loops
 * (BNN_SGDMC.py:131-137, BNN_BBB.py:143-149, BNN_Dropout.py:78-84,
 * BNN_CDropout.py:158-164, BNN_SVI.py:70-76) and, when mean/var are given,
 * `preds.mean(0), preds.var(0)` of BNN.predict_mv (BNN.py:34-37), in one
 * launch; pred[S,N,nout] is only written when asked for.                   */
int32_t bnn_forward(const BnnForwardArgs* a, void* stream);
/* 1 if the tensor-core (tcgen05) forward takes this shape in this precision mode, else 0 with the reason in
 * bnn_last_error().  Per-sample banks and the dropout family (one shared network + per-sample masks, external or
 * in-kernel) are both supported.  BNN_PREC_FP32 takes every shape. */
int32_t bnn_forward_tc_supported(const BnnForwardArgs* a);
/* debug: timeline of the last tensor-core launch made with BNN_TC_TRACE=1 in the environment (5*4096*2 uint64) */
int32_t bnn_tc_trace_dump(unsigned long long* host_out);
/* tcgen05 self-test: D[128, N] = A[128, K] * B[N, K]^T (row-major device buffers, K % 8 == 0, K <= 64,
 * N % 16 == 0, N <= 256, passes = 1 (TF32) or 3 (3xTF32), + 16 to stage the A operand in tensor memory instead of
 * shared memory) through the kernels' own operand layout */
/* debug: issue-rate probe of the tensor pipe -- `iters` back-to-back M = 128 MMAs (mode 0 tf32 SS, 1 tf32 TS, 2 bf16 SS,
 * 3 bf16 TS, 4 mixed stage pattern, 5 3xTF32 stage pattern); dev_out_clk[0] = issue cycles, [1] = cycles to completion */
int32_t bnn_tc_mma_bench(int32_t mode, int32_t N, int32_t iters, long long* dev_out_clk, void* stream);
int32_t bnn_tc_selftest(const float* A, const float* B, float* D, int32_t K, int32_t N, int32_t passes, void* stream);

/* ---- (e) standalone reduction and metrics ------------------------------- */
/* replaces: BNN.predict_mv's `preds.mean(dim=0), preds.var(dim=0)` (BNN.py:37)
 * on an already materialised pred[S, M] (M = N*nout).                       */
int32_t bnn_moments(const float* pred, int64_t S, int64_t M, float* mean, float* var,
                    double* part_mean, double* part_m2, void* workspace, int64_t workspace_bytes,
                    void* stream);
int64_t bnn_moments_workspace_bytes(int64_t S, int64_t M);

/* Chan merge of R shards' (count, mean, M2) -> mean/var (unbiased).  Used after
 * the NCCL all-gather/all-reduce of the per-point moments.
 * counts_host is a HOST array of R sample counts.                           */
int32_t bnn_moments_merge(const double* part_mean, const double* part_m2, const int64_t* counts_host,
                          int32_t R, int64_t M, float* mean, float* var, void* stream);

/* same merge with a float64 (mean, M2) result: several chunks of one shard -> the shard's moments */
int32_t bnn_moments_merge_partials(const double* part_mean, const double* part_m2, const int64_t* counts_host,
                                   int32_t R, int64_t M, double* out_mean, double* out_m2, void* stream);

/* ---- (e) multi-GPU: exchange of the per-point moments over NCCL ----------------- */
/* One process per GPU, samples sharded over the ranks: every rank passes the float64 (mean, M2) of ITS `count`
 * samples for all M points; float64 raw moments (n mean, M2 + n mean^2) are reduce-scattered (ncclSum), each rank finalises
 * M / R points, and with gather != 0 an all-gather hands every rank the full float32 mean / unbiased var (what
 * BNN.predict_mv returns, BNN.py:34-37); with gather == 0 a rank only writes its own slice [rank*ceil(M/R), ...) of mean / var.
 * Payload per rank: 16 M bytes reduce-scatter (+ 8 M all-gather), independent of R; stream-ordered, no host sync.
 * nccl_comm is an ncclComm_t created by the caller from the SAME libnccl that bnn_nccl_load opened (NULL path =
 * "libnccl.so.2" on the default search path; call it first with the path of the NCCL your process uses).               */
int32_t bnn_nccl_load(const char* path);
int64_t bnn_moments_allreduce_workspace_bytes(int64_t M, int32_t n_ranks);
int32_t bnn_moments_allreduce(void* nccl_comm, const double* part_mean, const double* part_m2, int64_t count, int64_t M,
                              float* mean, float* var, int32_t gather, void* workspace, int64_t workspace_bytes, void* stream);

/* replaces: BNN.validate's rmse / nll lines (BNN.py:30-31) and, with
 * noise_var != NULL (per output, device), experiments/SGDMC.py:83-93.
 * out = {rmse[nout], nll[nout]} floats followed by one float holding the count
 * of non-positive variances (the reference raises there: Normal(scale<=0)).
 * scratch: 2*nout+1 doubles of device memory owned by the caller (the FP64 accumulators).  */
int32_t bnn_metrics(const float* mean, const float* var, const float* y, const float* noise_var,
                    int64_t N, int32_t nout, float* out /* [2*nout+1] */, double* scratch /* [2*nout+1], device */,
                    void* stream);

/* ---- (b) reparameterised weight draws + KL ------------------------------ */
enum { BNN_SCALE_BBB = 0 /* sqrt(softplus(max(rho,-35))) BNN_BBB.py:24-27 */,
       BNN_SCALE_SVI = 1 /* softplus(rho)               BNN_SVI.py:42-46 */ };
/* replaces: GaussianLinear.rsample/sample_linear (BNN_BBB.py:24-27,44-51) for all
 * layers and all S draws at once, the guide draw of BNN_SVI.py:40-48, and the
 * KL sum of BNN_BBB.loss (BNN_BBB.py:107-110; kl_out may be NULL).
 * mu, rho: packed [stride]; eps: [S, stride] or NULL (= Philox in-kernel,
 * stream (seed, sample_offset + s)); W_out: bank [S, stride].                */
int32_t bnn_reparam_kl(const BnnMlpDesc* d, int32_t scale_mode, const float* mu, const float* rho,
                       const float* eps, uint64_t seed, int64_t sample_offset, int64_t S,
                       float* W_out, float weight_std, double* kl_out, void* stream);

/* ---- (d) dropout masks --------------------------------------------------- */
/* replaces: the mask draws of BNN_Dropout.py:70-75 (kind BNN_MASK_BERNOULLI) and
 * CDropoutLinear.sample BNN_CDropout.py:59-65 (BNN_MASK_CONCRETE); same Philox
 * streams as the in-kernel path of bnn_forward, so the two agree bit for bit.
 * uniforms (optional, [S, mask_len]) replaces Philox by caller-supplied u.     */
int32_t bnn_dropout_masks(const BnnMlpDesc* d, const BnnMaskSpec* spec, const float* uniforms,
                          int64_t S, float* masks_out /* [S, mask_len] */, void* stream);
/* materialise W[s] = base * diag(mask[s]) per masked layer (the reference's own
 * arithmetic order, BNN_Dropout.py:74 / BNN_CDropout.py:64)                    */
int32_t bnn_apply_masks(const BnnMlpDesc* d, const int32_t* layer_masked, const float* base,
                        const float* masks, int64_t mask_stride, int64_t S, float* W_out, void* stream);

/* ---- (c) pSGLD update ---------------------------------------------------- */
/* replaces: pSGLD.step() + the LambdaLR factor (BNN_SGDMC.py:86-87,100-101).
 * theta/grad/V: [n]; elements [0, n_head) use lr_head, the rest lr_tail (the two
 * param groups of BNN_SGDMC.py:96-98).  noise: [n] or NULL (= Philox, stream
 * (seed, step)).  V <- a V + (1-a) g^2; G = 1/(lam + sqrt(V));
 * theta <- theta - (0.5 lr G g + sqrt(lr G) xi).  UNPINNED: pybnn absent.      */
int32_t bnn_psgld_step(float* theta, const float* grad, float* V, int64_t n, int64_t n_head,
                       float lr_head, float lr_tail, float alpha, float lam, const float* noise,
                       uint64_t seed, int64_t step, void* stream);

/* ---- (f1) fused SGLD step ---------------------------------------------------- */
/* replaces: the body of BNN_SGDMC.sgld_steps (BNN_SGDMC.py:76-91) -- shuffled minibatch (:80,110), log_prior / log_lik /
 * loss (:61-74,83), loss.backward() (:85), pSGLD.step() and the LambdaLR factor (:86-87,96-101) -- for n_steps
 * consecutive steps in ONE persistent cooperative kernel launch (no library GEMM, no elementwise launches; see
 * csrc/sgld_step.cu).  The chain state is owned by the caller:
 *   theta  [2][bnn_sgld_theta_len]  packed network | log_precs[nout] | pad, double-buffered by step parity;
 *   thetaT [2][bnn_sgld_theta_t_len] transposed weight copies of Linear 1.. (maintained by the kernel);
 *   V      [bnn_sgld_theta_len]     RMSprop preconditioner state (pybnn initialises it to ones);
 *   workspace                       bnn_sgld_workspace_bytes() bytes of activations / output gradients / counters.
 * bnn_sgld_init: after writing the initial state into theta[cur] (and V), zeroes the workspace and builds thetaT[cur].
 * bnn_sgld_run : steps t0 .. t0+n_steps-1 on the batches perm[perm_off + i*batch ..] (the last batch of an epoch may be
 *   short; a launch must not run past num_train).  Reads theta[cur]; after the call the current buffer is
 *   cur ^ (n_steps & 1).  loss_out (device, may be NULL) receives U = -(loglik * N/|b| + logprior) of the last step.
 *   grad_out (device [theta_len], may be NULL): gradient-only mode for tests -- ONE step, writes dU/dtheta, no update.
 * Update rule (UNPINNED, pybnn absent): V <- a V + (1-a) g^2; G = 1/(lam + sqrt(V)) (precond_mode 0) or 1/sqrt(V + lam)
 *   (precond_mode 1); theta <- theta - (0.5 lr_t G g + sqrt(lr_t G) xi), lr_t = lr * float32((1+t)^-0.33), xi = Philox
 *   (seed, t); layout padding is never touched.                                                                        */
typedef struct BnnSgldChain {
  BnnMlpDesc desc;
  int32_t batch;             /* rows per minibatch (BNN_SGDMC.py:30) */
  float* theta;
  float* thetaT;
  float* V;
  void* workspace;
  int64_t workspace_bytes;
  const float* X;            /* [num_train, dims[0]] */
  const float* y;            /* [num_train, nout]    */
  const int64_t* perm;       /* epoch permutation [num_train] (DataLoader(shuffle=True), :110) */
  int64_t num_train;
  float lr_weight, lr_noise; /* the two param groups of :96-98 */
  float precond_decay;       /* a   (pybnn default 0.99) */
  float diagonal_bias;       /* lam (pybnn default 1e-5) */
  float alpha_n, beta_n;     /* Gamma prior of the precisions (:35-36, :47) */
  float gain;                /* weight-prior gain (5/3, :51) */
  int32_t precond_mode;
  uint64_t seed;
} BnnSgldChain;
int64_t bnn_sgld_theta_len(const BnnMlpDesc* d);
int64_t bnn_sgld_theta_t_len(const BnnMlpDesc* d);
int64_t bnn_sgld_workspace_bytes(const BnnMlpDesc* d, int32_t batch);
int32_t bnn_sgld_init(const BnnSgldChain* c, int32_t cur, void* stream);
int32_t bnn_sgld_run(const BnnSgldChain* c, int32_t cur, int64_t t0, int64_t perm_off, int32_t n_steps, float* loss_out,
                     float* grad_out, void* stream);
/* Several chains on one GPU: a chain's step is a dependency chain of seven tile stages, so one chain leaves SMs idle about a
 * third of the time.  Capping the CTAs of a launch (process-wide; 0 = no cap) and running each chain's launches on its own stream
 * lets the chains interleave: cfg 5 on one B200, 1 x 148 CTAs 1.67e4 steps/s, 2 x 74 2.28e4, 4 x 37 2.42e4 in total
 * (tools/sgld_multi_chain.py).  Launches are cooperative, so a chain starts only when all of its CTAs fit beside the others'. */
int32_t bnn_sgld_set_grid_cap(int32_t ctas);
/* entries [offset, offset + count) of the pseudo-random permutation of [0, N) keyed by (seed, epoch): the shuffled order of
 * DataLoader(shuffle=True) (BNN_SGDMC.py:110) without materialising all N entries -- a chain only consumes a prefix */
int32_t bnn_sgld_perm(uint64_t seed, uint32_t epoch, int64_t N, int64_t offset, int64_t count, int64_t* out, void* stream);
/* debug: per-CTA job timeline of the last two steps of the following bnn_sgld_run launches: [grid][recs][12] words
 * (see tools/sgld_trace.py); *recs_out = records per CTA */
int32_t bnn_sgld_trace(int32_t enable, long long* host_out, int32_t* grid_out, int32_t* recs_out);
/* synchronises; *status_host = 0 ok, 1 = a dependency wait of the last launch timed out (chain state undefined) */
int32_t bnn_sgld_status(const BnnSgldChain* c, int32_t* status_host, void* stream);

/* ---- (f2) fused Bayes-by-Backprop training step ----------------------------------- */
/* replaces: the minibatch loop body of BNN_BBB.train (BNN_BBB.py:121-126) -- GaussianLinear.forward with local
 * reparameterisation (:29-37), MSELoss (:104-106), the KL term (:107-110), loss.backward() and Adam.step() -- for n_steps
 * consecutive steps in ONE single-CTA persistent kernel (csrc/bbb_train.cu), for networks whose variational parameters, Adam
 * moments and activations fit one SM's shared memory (bnn_bbb_train_supported; the reference's 1-50-50-1 at batch 32 does).
 * mu / rho: packed bias-last vectors [stride] (the layout of bnn_reparam_kl), updated in place; adam: [4][stride] = first / second
 * moments of mu, then of rho (zero before the first step); t0 = Adam steps taken so far (bias correction).
 * Batches: rows perm[perm_off + i * batch ..] of X / y, the last one of an epoch may be short.
 * mse_out / kl_out ([n_steps] device floats or NULL): the step's MSE and KL(q || N(0, weight_std)) before its update.
 * eps_in (tests): the local-reparameterisation noise, [n_steps][bnn_bbb_train_eps_per_step] = per layer [batch][out]; NULL =
 * in-kernel Philox (seed, t0 + i).                                                                                        */
typedef struct BnnBbbTrain {
  BnnMlpDesc desc;
  int32_t batch;
  float* mu;
  float* rho;
  float* adam;
  const float* X;            /* [num_train, dims[0]] */
  const float* y;            /* [num_train, nout]    */
  const int64_t* perm;       /* epoch permutation [num_train] */
  int64_t num_train;
  float lr, beta1, beta2, eps;   /* torch.optim.Adam: 1e-2 (BNN_BBB.py:93), 0.9, 0.999, 1e-8 */
  float kl_factor;           /* BNN_BBB.py:95 */
  float weight_std;          /* prior std, BNN_BBB.py:94 */
  uint64_t seed;
} BnnBbbTrain;
int32_t bnn_bbb_train_supported(const BnnMlpDesc* d, int32_t batch);
int64_t bnn_bbb_train_eps_per_step(const BnnMlpDesc* d, int32_t batch);
int32_t bnn_bbb_train_run(const BnnBbbTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, float* mse_out, float* kl_out,
                          const float* eps_in, void* stream);

/* replaces: the minibatch loop body of BNN_Dropout.train (BNN_Dropout.py:57-62) -- NN_Dropout.forward with fresh Bernoulli keep
 * masks on the inputs of every Linear wider than one unit (:15-21), MSELoss(reduction='sum') (:55,59), loss.backward() and
 * Adam.step() with weight decay l2_reg on the weights only (:43-52) -- for n_steps consecutive steps in ONE single-CTA persistent
 * kernel (csrc/dropout_train.cu), for networks that fit one SM's shared memory (bnn_dropout_train_supported; the reference's
 * 9-100-100-1 at batch 32 does).  W: the packed network [stride] (the layout of the forward path), updated in place;
 * adam: [2][stride] first / second moments (zero before the first step); t0 = Adam steps taken so far.
 * loss_out ([n_steps] device floats or NULL): the step's sum of squared errors before its update (epoch sum -> noise_level, :63).
 * mask_in (tests): keep factors [n_steps][bnn_dropout_train_mask_per_step] = per masked layer [batch][in] (each block padded to a
 * multiple of 4 floats); NULL = in-kernel Philox (seed, t0 + i).                                                             */
typedef struct BnnDropoutTrain {
  BnnMlpDesc desc;
  int32_t batch;
  float* W;
  float* adam;
  const float* X;            /* [num_train, dims[0]] */
  const float* y;            /* [num_train, nout]    */
  const int64_t* perm;       /* epoch permutation [num_train] */
  int64_t num_train;
  float lr, beta1, beta2, eps;   /* torch.optim.Adam: 1e-2 (BNN_Dropout.py:35), 0.9, 0.999, 1e-8 */
  float l2_reg;              /* BNN_Dropout.py:34 */
  float p_drop;              /* BNN_Dropout.py:32 */
  uint64_t seed;
} BnnDropoutTrain;
int32_t bnn_dropout_train_supported(const BnnMlpDesc* d, int32_t batch);
int64_t bnn_dropout_train_mask_per_step(const BnnMlpDesc* d, int32_t batch);
int32_t bnn_dropout_train_run(const BnnDropoutTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, float* loss_out,
                              const float* mask_in, void* stream);

/* replaces: the minibatch loop body of BNN_CDropout.train (BNN_CDropout.py:121-133) -- CDropout.forward's relaxed keep factors on
 * the input of EVERY Linear (:21-34; temperature 0.1, one learnable p_logit per layer, :13-19), MSELoss(reduction='sum'), the
 * regulariser (wreg - ent) * rows / num_train of BNN_CDropout.reg (:141-151) + CDropoutLinear.reg (:53-57), loss.backward() and
 * Adam.step() on W, b and p_logit (:116) -- in the same single-CTA persistent kernel as bnn_dropout_train_run
 * (csrc/dropout_train.cu, CONCRETE = true).  p_logit: [n_linear] device floats, p_adam: [2][n_linear] (zero before the first step).
 * noise_level: ONE device float, read at launch (the regulariser's noise_level^2, :144); with epoch_end != 0 the launch ends by
 * storing max(min_noise, sqrt(sum of its squared errors / num_train)) (:134), so an epoch needs no host round trip.
 * stats_out ([n_steps][3] device floats or NULL): squared errors, wreg * rows / N, ent * rows / N of each step (:127-132).
 * u_in (tests): the uniforms of torch.rand_like [n_steps][bnn_cdropout_train_u_per_step] = per layer [batch][in] (each block
 * padded to a multiple of 4 floats); NULL = in-kernel Philox (seed, t0 + i).                                                    */
typedef struct BnnCDropoutTrain {
  BnnMlpDesc desc;
  int32_t batch;
  float* W;
  float* adam;
  float* p_logit;
  float* p_adam;
  float* noise_level;
  const float* X;            /* [num_train, dims[0]] */
  const float* y;            /* [num_train, nout]    */
  const int64_t* perm;       /* epoch permutation [num_train] */
  int64_t num_train;
  float lr, beta1, beta2, eps;   /* torch.optim.Adam: 1e-2 (BNN_CDropout.py:107), 0.9, 0.999, 1e-8 */
  float lscale, dr;          /* BNN_CDropout.py:108-109 */
  float min_noise;           /* BNN_CDropout.py:105 */
  uint64_t seed;
} BnnCDropoutTrain;
int32_t bnn_cdropout_train_supported(const BnnMlpDesc* d, int32_t batch);
int64_t bnn_cdropout_train_u_per_step(const BnnMlpDesc* d, int32_t batch);
int32_t bnn_cdropout_train_run(const BnnCDropoutTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, int32_t epoch_end,
                               float* stats_out, const float* u_in, void* stream);

/* raw Philox stream dump for tests: out[n] 32-bit words / normals */
int32_t bnn_philox_words(uint64_t seed, uint32_t subsequence, uint32_t tag, int64_t n, uint32_t* out, void* stream);
int32_t bnn_philox_normals(uint64_t seed, uint32_t subsequence, uint32_t tag, int64_t n, float* out, void* stream);

/* measurement helper: sustained FP32 FFMA TFLOP/s of the current device (dependent-chain
 * microbenchmark, best of 4) -- the roofline denominator of the SIMT forward */
double bnn_measure_fp32_tflops(void);

/* ---- host-buffer convenience (end-to-end path) --------------------------- */
/* Same as bnn_forward but X / W / mask and all outputs are HOST pointers
 * (pinned for full PCIe speed); copies in, runs, copies out, synchronises.   */
int32_t bnn_forward_host(const BnnForwardArgs* a_host);
void bnn_release_workspace(void);

#ifdef __cplusplus
}
#endif
#endif /* BNN_B200_H_ */
