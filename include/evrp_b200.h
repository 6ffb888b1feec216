/* evrp_b200.h — C ABI of libevrp_b200.so (hand-written sm_100a CUDA behind plain pointers).
 *
 * Drop-in boundary for ONE hot path of mengchengTang/DRL-Energy-optimal-Routing-for-Electric-Vehicles:
 * the batched autoregressive rollout of the Transformer pointer policy and the EVRP environment step.
 * The reference is pure Python/PyTorch and has no FFI of its own; every entry point below cites the
 * reference interface (file:line under /root/reference/EM-EVRP) that it replaces.  A maintainer binds
 * these with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; tensors are dense, row-major,
 *    fp32 unless stated; node layout is [depot, F stations, N customers], S = 1+F+N <= EVRP_MAX_S
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing and keeps
 *    no state between calls; scratch comes from the caller (`*_workspace_bytes`)
 *  - return value: 0 ok; <0 = -(cudaError_t) of a CUDA runtime failure or EVRP_E_* argument error.
 *    Domain errors (the reference's Python asserts) are reported through a device-side int32 status word
 *    the caller passes in and reads back together with the results (bit mask EVRP_ST_*).
 *  - never throws, never falls back to the CPU.
 */
#ifndef EVRP_B200_H
#define EVRP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVRP_ABI_VERSION 22
#define EVRP_MAX_S 128          /* mask bits live in 4 x uint32 per instance                         */
#define EVRP_D 128              /* embedding_dim (trainer.py:261 / run.py --embedding_dim default)   */
#define EVRP_H 8                /* n_heads (nets/DRLModel.py:52)                                     */
#define EVRP_FF 512             /* feed_forward_hidden (nets/DRLModel.py:64, nets/GraphEncoder.py:159) */
#define EVRP_MAX_LAYERS 8

/* argument errors (returned negated) */
#define EVRP_E_BADARG 1000
#define EVRP_E_TOO_MANY_NODES 1001
#define EVRP_E_WORKSPACE 1002

/* device status bits (OR-ed into *status by kernels) */
#define EVRP_ST_NAN_LOGIT 1        /* nets/DRLModel.py:324,405  "Probs should not contain any nans"          */
#define EVRP_ST_INFEASIBLE_PICK 2  /* nets/DRLModel.py:328      greedy/forced action is masked               */
#define EVRP_ST_STEP_CAP 4         /* nets/DRLModel.py:255      an instance did not finish within max_steps  */
#define EVRP_ST_NONUNIFORM_DYNAMIC 8 /* load/SOC/time rows of `dynamic` are not one scalar broadcast over S   */
#define EVRP_ST_NO_FEASIBLE 16     /* every node masked for an unfinished instance                            */
#define EVRP_ST_STUCK_AT_DEPOT 32  /* an instance is back at the depot with open demand nothing can reach (rule 8 of
                                      problems/EVRP.py:220-221 leaves only the depot): the reference repeats (pi 0, ll 0,
                                      R 0) until its 1000-step cap (nets/DRLModel.py:255); the kernel retires the row at
                                      the fixed point, the outputs beyond are the zeros already there, and the caller
                                      reports T = the step cap                                                  */

/* Vehicle / problem constants.  problems/EVRP.py:21-41.  The caller fills the folded fp32 constants with
 * the SAME double-precision expressions the reference evaluates in Python, so that the state recipe is
 * bit-reproducible.  evrp_phys_default() does that in C. */
typedef struct {
  float aero;        /* fp32(0.5*Cd*A*Ad*(velocity/3.6)**2)             problems/EVRP.py:247           */
  float kd;          /* fp32(motor_d*battery_d)                          problems/EVRP.py:251           */
  float kr;          /* fp32(motor_r*battery_r)                          problems/EVRP.py:252           */
  float g, Cr, velocity, mc, w, max_load;
  float start_soc, t_limit, charging_time, serve_time;
  float charge_plus_serve;   /* fp32(charging_time + serve_time)         problems/EVRP.py:214           */
  float inv_velocity, inv_3600; /* fp32(1/x), used only when div_mode==1                                  */
  int32_t div_mode;  /* 0: x / scalar is an IEEE division (ATen CPU); 1: x * fp32(1/scalar) (ATen CUDA)  */
  int32_t objective; /* 0: EM, reward = energy (EM-EVRP/problems/EVRP.py:280); 1: DM, reward = distance
                        (DM-EVRP/problems/EVRP.py:231,280)                                               */
} evrp_phys;

void evrp_phys_default(evrp_phys* p, double velocity, double start_soc, double t_limit, double max_load);

int evrp_abi_version(void);
/* number of SMs / compute capability of the current device; <0 on error */
int evrp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * (1) stand-alone environment step:  VehicleRoutingDataset.update_dynamic  problems/EVRP.py:226-280
 *                                    VehicleRoutingDataset.update_mask     problems/EVRP.py:105-223
 * dynamic [B,4,S] (load, demand, SOC, time), distances/slope [B,S,S], now/chosen int64 [B].
 * ---------------------------------------------------------------------------------------------- */
int evrp_update_dynamic(const evrp_phys* phys, const float* dynamic, const float* distances, const float* slope,
                        const int64_t* now_idx, const int64_t* chosen_idx, int B, int S, int F,
                        float* new_dynamic /*[B,4,S]*/, float* energy /*[B]*/, float* distance /*[B]*/,
                        void* stream);

/* `any_demand` is a 1-int device scratch word (zeroed by the call); the all-zero mask the reference returns
 * when the whole batch is finished (problems/EVRP.py:127-128) is applied by a second tiny kernel. */
int evrp_update_mask(const evrp_phys* phys, const float* dynamic, const float* distances, const float* slope,
                     const int64_t* chosen_idx, int B, int S, int F, float* mask /*[B,S] 0/1*/,
                     int32_t* any_demand, void* stream);

/* distances / slope build of the 3-tuple input (static, dynamic, Elevations): nets/DRLModel.py:124-133 and the dataset's
 * own build problems/EVRP.py:87-92.  static_xy [B,2,S], elevations [B,S] -> distances, slope [B,S,S]. */
int evrp_pairwise_build(const float* static_xy, const float* elevations, int B, int S, float* distances, float* slope,
                        void* stream);

/* on-device instance generation: B synthetic instances with the distribution and value grid of the dataset
 * constructor problems/EVRP.py:43-92 (layout [depot, F stations, N customers]) from a counter-based generator: static_xy
 * [B,2,S], dynamic [B,4,S] = [1, demand, Start_SOC, t_limit], elevations [B,S]; S = 1 + F + N <= EVRP_MAX_S. */
int evrp_generate_instances(uint64_t seed, int B, int N, int F, float start_soc, float t_limit, int max_load, int max_demand,
                            float* static_xy, float* dynamic, float* elevations, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) encoder + decoder precompute:  AttentionModel._init_embed        nets/DRLModel.py:192-200
 *                                    GraphAttentionEncoder.forward     nets/GraphEncoder.py:202-214
 *                                    AttentionModel._precompute        nets/DRLModel.py:345-362
 * Weights arrive PACKED k-major ([in_features][out_features], out contiguous) so that every GEMM reads its
 * B operand coalesced; evrp_b200/model.py::pack_weights builds them from the reference's state_dict keys
 * (one transpose/concat per tensor, redone only when the parameters change).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float *Wqkv;             /* [128,384]  column h*16+d of block 0/1/2 = W_query/W_key/W_val[h,:,d]
                                               (embedder.layers.l.0.module.*, nets/GraphEncoder.py:41-43)  */
  const float *Wo;               /* [128,128]  W_out.view(H*16,128)          nets/GraphEncoder.py:104-107   */
  const float *bn1_scale, *bn1_shift;   /* [128] eval-mode BatchNorm1d folded to y = x*scale + shift
                                               (layers.l.1.normalizer, nets/GraphEncoder.py:142-145)        */
  const float *ff1_w, *ff1_b;    /* [128,512],[512]  layers.l.2.module.0.weight^T / bias                    */
  const float *ff2_w, *ff2_b;    /* [512,128],[128]  layers.l.2.module.2.weight^T / bias                    */
  const float *bn2_scale, *bn2_shift;   /* layers.l.3.normalizer                                            */
  /* tcgen05 path (gemm_mode 0): the same four weight matrices in nn.Linear layout [out][in] (in contiguous), each
     split into two TF32 planes with round-to-nearest, w = hi + lo (evrp_b200/policy_math.py::split_tf32)     */
  const float *Wqkv_hi, *Wqkv_lo;       /* [384,128] */
  const float *Wo_hi, *Wo_lo;           /* [128,128] */
  const float *ff1_hi, *ff1_lo;         /* [512,128] */
  const float *ff2_hi, *ff2_lo;         /* [128,512] */
} evrp_layer_weights;

typedef struct {
  const float *depot_w, *depot_b;      /* [128,2],[128]  init_embed_depot_and_station (state_dict layout)   */
  const float *station_w, *station_b;  /* [128,2],[128]  init_embed_station                                 */
  const float *cust_w, *cust_b;        /* [128,3],[128]  init_embed                                         */
  int32_t n_layers;
  evrp_layer_weights layers[EVRP_MAX_LAYERS];
  const float *node_proj;              /* [128,384]  project_node_embeddings.weight^T; columns 0..127 glimpse-K,
                                          128..255 glimpse-V, 256..383 logit-K with project_out folded in:
                                          lK_fold = project_out.weight^T @ W_logitK, so that
                                          logits[j] = heads . lK[j]           (nets/DRLModel.py:436-444)     */
  const float *fixed_ctx_w;            /* [128,128]  project_fixed_context.weight^T                         */
  const float *node_proj_hi, *node_proj_lo;   /* [384,128] TF32 planes of node_proj^T (tcgen05 path)         */
  /* nets/AM.py policy variant only, else NULL: project_step_context.weight[:, :128] (nets/AM.py:90),
     the part of the step-context projection that multiplies emb[now] (nets/AM.py:336-337,352-377).  The encoder
     tabulates it per node (step_tab below), so the rollout reads one 512-byte row per step instead of a mat-vec. */
  const float *step_proj;                     /* [128,128] k-major (weight[:, :128]^T), gemm_mode 1            */
  const float *step_proj_hi, *step_proj_lo;   /* [128,128] [out][in] TF32 planes, gemm_mode 0                  */
  int32_t gemm_mode;                   /* 2: ONE fused persistent tcgen05 kernel for the whole encoder (fp16-pair split
                                          operands, fp32-faithful; needs fused_w / fused_params below); 0: one tcgen05
                                          3xTF32 GEMM / attention launch per sub-layer; 1: fp32 FFMA GEMMs            */
  int32_t attn_mode;                   /* self-attention core (QK^T, softmax, PV): 0: tcgen05 3xTF32, probabilities as
                                          the TMEM A operand (used when gemm_mode == 0); 1: fp32 FFMA2 kernel          */
  /* fused encoder kernel (gemm_mode 2, csrc/evrp_encoder_fused.cu): every weight matrix of the encoder and of the node
     projection as fp16 (hi, lo) operand blocks [128 out][64 in] in the kernel's consumption order, each matrix scaled
     by a power of two (evrp_b200/policy_math.py::pack_fused_encoder documents the order):
       per layer: W_qkv chunks Q, K, V x (K block 0 hi, lo, K block 1 hi, lo); W_out; then FF1 chunk 0, FF1 chunk 1,
       FF2 K blocks 0-1, FF1 chunk 2, FF2 K blocks 2-3, FF1 chunk 3, FF2 K blocks 4-5, FF2 K blocks 6-7;
       then the node projection chunks (glimpse K, glimpse V, folded logit K[, AM step table])                       */
  const void *fused_w;                 /* fp16 [n_layers * 48 + fused_node_chunks * 4][128][64]                        */
  const float *fused_params;           /* fp32 [n_layers][1280]: bn1 scale 128 | bn1 shift 128 | ff1 bias 512 | ff2 bias
                                          128 | bn2 scale 128 | bn2 shift 128 | 1/scale of W_qkv, W_out, FF1, FF2 | pad;
                                          then [1]: 1/scale of the node projection                                   */
  int32_t fused_node_chunks;           /* 3, or 4 when the AM step table is packed behind the node projection         */
} evrp_encoder_weights;

/* scratch of the per-sub-layer paths (gemm_mode 0 / 1); the fused kernel (gemm_mode 2) needs none (workspace may be NULL) */
size_t evrp_encoder_workspace_bytes(int B, int S);

/* Eval-mode forward (BatchNorm folded to per-channel affine).  Outputs: emb [B,S,128] node embeddings,
 * kvl [B,S,384] = per node (glimpse-K | glimpse-V | folded logit-K), fixed_ctx [B,128], and -- AM variant only,
 * else NULL -- step_tab [B,S,128] = emb @ project_step_context.weight[:, :128]^T. */
int evrp_encoder_forward(const evrp_encoder_weights* w, const float* static_xy /*[B,2,S]*/,
                         const float* dynamic /*[B,4,S]*/, int B, int S, int F,
                         float* emb, float* kvl, float* fixed_ctx, float* step_tab,
                         void* workspace, size_t workspace_bytes, void* stream);

/* diagnostics of the fused encoder kernel (synchronises): returns 1 and fills out8_host = {code of the first barrier
 * wait that did not complete within ~2 s, CTA, thread, parity, ...} if a launch since the last reset aborted, else 0 */
int evrp_encoder_fused_debug(int* out8_host, int reset);

/* ------------------------------------------------------------------------------------------------
 * (3) persistent rollout:  AttentionModel._inner        nets/DRLModel.py:240-280
 *                          route_feature                :203-237
 *                          _get_log_p/_one_to_many_logits :379-452
 *                          _select_node                 :316-342
 *                          _calc_log_likelihood         :182-190
 *                          + update_dynamic / update_mask fused (problems/EVRP.py:105-280)
 * One launch for the whole rollout; instances retire individually; outputs beyond an instance's last
 * step keep the zeros the caller memset (the reference pads finished rows with pi=0, ll=0, R=0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float *w1_veh;   /* [3,512]    (FF_tour.0.weight[:, :128] @ tour.weight)^T, folded in fp64     */
  const float *w1_max;   /* [128,512]  FF_tour.0.weight[:, 128:]^T                                      */
  const float *b1;       /* [512]      FF_tour.0.bias                                                   */
  const float *w2;       /* [512,128]  FF_tour.2.weight^T                                               */
  const float *b2;       /* [128]      FF_tour.2.bias                                                   */
  /* tcgen05 path (ff_mode 0): the two FF_tour matrices in nn.Linear layout [out][in], each scaled by a power of two
     (exact) and split into two fp16 planes with round-to-nearest, s*w = hi + lo: 22 significant bits, the same as a
     TF32 pair at half the bytes (evrp_b200/policy_math.py::split_f16); streamed by TMA every decode step           */
  const void *w1max_hi, *w1max_lo;    /* fp16 [512,128]  w1_scale * FF_tour.0.weight[:, 128:]                      */
  const void *w2_hi, *w2_lo;          /* fp16 [128,512]  w2_scale * FF_tour.2.weight                               */
  const float *am_w_load;             /* [128] project_step_context.weight[:, 128] (the load column), policy 1     */
  float w1_inv_scale, w2_inv_scale;   /* 1 / scale (powers of two), applied to the fp32 accumulators               */
  int32_t ff_mode;       /* 0: FF_tour on tcgen05 tensor cores (3 x fp16 products, fp32-faithful), one 576-thread CTA
                            per SM; 1: fp32 FFMA inside a 256-thread CTA (2 per SM)                              */
  int32_t policy;        /* step query of the decoder.  0: nets/DRLModel.py:394  fixed_ctx + FF_tour(route context) +
                            emb[now] (the fields above).  1: nets/AM.py:336-337  fixed_ctx + project_step_context(
                            [emb[now], load]) = fixed_ctx + step_tab[now] + load * am_w_load: no FF_tour, and the `emb`
                            argument of evrp_rollout is the encoder's step_tab                                   */
  const float *inv_scales; /* optional DEVICE [2] = {w1_inv_scale, w2_inv_scale}: when set the kernel reads the scales here
                              (evrp_pack_decoder derives them on the device, so packing needs no host read-back)       */
  const void *ff_stream;   /* optional: the same four fp16 planes as ONE 512 KB stream in the order the rollout kernel consumes
                              them: 16 stages x (hi 16 KB | lo 16 KB), stage i < 8 = rows 128(i/2).. of W1[:, 128:], K block
                              i%2 (64 inputs); stage 8 + k = W2, K block k; every 16 KB plane is a [128 rows][128 B] tile with
                              the 128-byte shared-memory swizzle already applied (16-byte chunk c of row r sits at c ^ (r & 7)).
                              When set, a stage is ONE contiguous 32 KB bulk copy instead of 256 strided 128-byte rows        */
} evrp_decoder_weights;

/* Decoder half of the weight packing in two kernels and without a host read-back (a REINFORCE step re-packs it after every
 * optimizer step): from tour.weight [128,3], FF_tour.0.weight [512,256] / bias, FF_tour.2.weight [128,512] (nn.Linear
 * layouts, nets/DRLModel.py:89-92) to w1_veh [3,512] = (W1[:, :128] tour.weight)^T folded in fp64, w1_max [128,512] and
 * w2 [512,128] k-major, the fp16 (hi, lo) planes of W1[:, 128:] and W2 scaled by a power of two chosen from the
 * largest magnitude (max |s w| ~ 2^12, as evrp_b200/policy_math.py::split_f16), inv_scales [2] = 1 / s, and the stream form
 * of the planes (ff_stream).
 * `scratch` is 2 device floats (zeroed by the call). */
int evrp_pack_decoder(const float* tour_w, const float* w1, const float* w2, float* w1_veh, float* w1_max, float* w2_t,
                      void* w1max_hi, void* w1max_lo, void* w2_hi, void* w2_lo, float* inv_scales, float* scratch,
                      void* ff_stream /* 512 KB, nullable: evrp_decoder_weights.ff_stream */, void* stream);

#define EVRP_DECODE_GREEDY 0
#define EVRP_DECODE_SAMPLE 1

typedef struct {
  int32_t decode_type;       /* EVRP_DECODE_*                          set_decode_type nets/DRLModel.py:104 */
  float   temp;              /* softmax temperature                    nets/DRLModel.py:403                 */
  float   tanh_clipping;     /* 10.0                                   nets/DRLModel.py:447                 */
  int32_t max_steps;         /* output width Tcap (<= 1000)            nets/DRLModel.py:255                 */
  uint64_t seed;             /* counter-based RNG seed for sampling                                         */
  int32_t n_replicas;        /* sample_many: rollouts per instance sharing one encoder output (>=1);
                                rollout row r uses instance r % B (replica-major, nets/DRLModel.py:291-294)   */
  /* 3-tuple input (static, dynamic, Elevations), nets/DRLModel.py:124-133: when evrp_rollout is called with
     distances == slope == NULL the kernel recomputes D[i,j] = sqrt((x_i-x_j)^2 + (y_i-y_j)^2) and G[i,j] = clamp((E_i -
     E_j) / (D + 1e-6), +-0.1) from these rows where it needs them (same separately rounded ops as evrp_pairwise_build,
     bit-identical results): no [B,S,S] tensor exists at all                                                        */
  const float *static_xy;    /* [B,2,S] or NULL                                                              */
  const float *elevations;   /* [B,S] or NULL                                                                */
} evrp_rollout_cfg;

int evrp_rollout(const evrp_phys* phys, const evrp_decoder_weights* w, const evrp_rollout_cfg* cfg,
                 const float* emb /*[B,S,128]*/, const float* kvl /*[B,S,384]*/, const float* fixed_ctx /*[B,128]*/,
                 const float* dynamic0 /*[B,4,S]*/, const float* distances /*[B,S,S] or NULL*/, const float* slope /*[B,S,S] or NULL*/,
                 int B, int S, int F,
                 const int64_t* forced_actions /*[B*R,Tcap] or NULL*/, const float* uniforms /*[B*R,Tcap] or NULL*/,
                 int64_t* pi /*[B*R,Tcap]*/, float* ll /*[B*R,Tcap]*/, float* reward /*[B*R,Tcap]*/,
                 float* energy /*[B*R,Tcap] or NULL*/,
                 uint32_t* saved_mask /*[B*R,Tcap,4] or NULL*/, float* saved_veh /*[B*R,Tcap,3] or NULL*/,
                 int32_t* steps_used /*[B*R]*/, int32_t* t_max /*[1] atomicMax*/, int32_t* status /*[1]*/,
                 void* workspace /* evrp_rollout_workspace_bytes(B,S,R): rule-7 station constants per instance,
                                    demand row and exact running-max row per rollout row (L2-resident state) */,
                 size_t workspace_bytes, void* stream);
size_t evrp_rollout_workspace_bytes(int B, int S, int n_replicas);

/* sample_many reduction (nets/DRLModel.py:301-311): per instance min cost over replicas + its tour. */
int evrp_sample_min(const float* reward /*[B*R,Tcap]*/, const int64_t* pi, int B, int R, int Tcap,
                    float* mincost /*[B]*/, int32_t* argmin /*[B]*/, int64_t* minpi /*[B,Tcap]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) fused REINFORCE loss / advantage (trainer.py:116-121):  reward = R.sum(1); advantage = reward - baseline;
 *     loss = sum_b(advantage * ll.sum(1)) * inv_global_batch;  grad_ll[b,t] = advantage[b] * inv_global_batch.
 *     `baseline` is [B] (rollout baseline) or one scalar (exponential baseline); `scratch` is [B] floats.
 * ---------------------------------------------------------------------------------------------- */
int evrp_reinforce_loss(const float* ll /*[B,T]*/, const float* R /*[B,T]*/, const float* baseline,
                        int baseline_is_scalar, int B, int T, float inv_global_batch,
                        const float* global_batch_dev /* nullable: device scalar = global batch size (e.g. all-reduced over
                                                         ranks); when set, 1 / it replaces inv_global_batch (no host read) */,
                        float* reward /*[B]*/, float* advantage /*[B]*/, float* grad_ll /*[B,T]*/, float* loss /*[1]*/,
                        float* scratch /*[B]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) teacher-forced decoder log-probabilities and their backward (the differentiable half of
 *     nets/DRLModel.py:379-452 for the REINFORCE step, trainer.py:115-127).  Inputs: the step queries
 *     q[b,t,:] = fixed_ctx + route context + emb[now] (nets/DRLModel.py:394), the encoder's kvl records, and the
 *     rollout's trace (mask bit words, actions, step counts).  forward: ll[b,t] = log_softmax(clip * tanh(
 *     glimpse(q) . lK / sqrt(128)) / temp)[pi[b,t]], 0 on padded steps.  backward: d ll -> d q, d kvl.
 *     No [B,H,T,S] tensor is materialised; one CTA per instance, d kvl accumulated in shared memory.
 * ---------------------------------------------------------------------------------------------- */
/* `saved_stats` (nullable): [B*T, evrp_pointer_logp_saved_floats()] floats where the forward keeps the per-step heads /
 * softmax maxima and sums / log-sum-exp; handed to the backward it replaces the recompute passes over the records. */
int evrp_pointer_logp_saved_floats(void);
int evrp_pointer_logp_forward(const float* q /*[B,T,128]*/, const float* kvl /*[B,S,384]*/,
                              const uint32_t* mask_bits /*[B,T,4]*/, const int64_t* pi /*[B,T]*/,
                              const int32_t* steps /*[B]*/, int B, int T, int S, float tanh_clipping, float temp,
                              float* ll /*[B,T]*/, float* saved_stats, void* stream);
int evrp_pointer_logp_backward(const float* q, const float* kvl, const uint32_t* mask_bits, const int64_t* pi,
                               const int32_t* steps, const float* grad_ll /*[B,T]*/, int B, int T, int S,
                               float tanh_clipping, float temp, float* grad_q /*[B,T,128]*/,
                               float* grad_kvl /*[B,S,384]*/, const float* saved_stats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (6) encoder self-attention for the TRAINING path (nets/GraphEncoder.py:81-102 and its autograd backward): fp32 SIMT,
 *     flash-style (no [B,H,S,S] tensor).  qkv [B,S,384] = Q | K | V projections, head-major inside each block;
 *     heads [B,S,128]; lse [B,8,S] = row log-sum-exp of the scaled scores, kept for the backward.
 * ---------------------------------------------------------------------------------------------- */
int evrp_self_attention_forward(const float* qkv, int B, int S, float* heads, float* lse, void* stream);
int evrp_self_attention_backward(const float* qkv, const float* heads, const float* lse, const float* grad_heads,
                                 int B, int S, float* grad_qkv, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (7) the encoder's tcgen05 GEMM on its own (training path):  C[M,N] = epilogue( A[M,K] @ W[N,K]^T ), fp32-faithful
 *     3xTF32 (W given as TF32 hi / lo planes in nn.Linear layout [out][in]).  N % 128 == 0.  epilogue bits: 1 bias,
 *     2 ReLU, 4 residual, 8 per-column affine, 16 auxiliary rank-3 term, 32 gate (result kept where resid > 0: the
 *     gradient through a ReLU whose output is `resid`); supported: 0 (any K % 32 == 0), 1|2, 4|8, 1|2|16 and 32
 *     (K == 128), 1|4|8 (K == 512):  C = ((A W^T + bias + aux aux_w) [relu] + resid) * scale + shift with
 *     aux [M,3] row-major and aux_w [3][N] (the vehicle-scalar term of FF_tour.0, nets/DRLModel.py:215-235).
 *     evrp_split_tf32 builds the planes from a weight with arbitrary strides (a transposed view for the input-gradient
 *     product): hi = rna_tf32(w), lo = rna_tf32(w - hi), dense [rows][cols].
 * ---------------------------------------------------------------------------------------------- */
int evrp_gemm_3xtf32(int epilogue, const float* A, int lda, const float* W_hi, const float* W_lo, float* C, int ldc,
                     int M, int N, int K, const float* bias, const float* resid, int ldr, const float* scale,
                     const float* shift, const float* aux, const float* aux_w, void* stream);
int evrp_split_tf32(const float* w, int rows, int cols, long long stride_r, long long stride_c, float* hi, float* lo,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * (8) weight gradients of the REINFORCE backward (trainer.py:122 loss.backward() through the nn.Linear / W_query /
 *     W_key / W_val / W_out products of nets/GraphEncoder.py:55-118,153-170 and FF_tour nets/DRLModel.py:233-237):
 *     gW [N_out, K_in] = gy[M, N_out]^T x[M, K_in], the contraction over all M = B*S (or B*T) rows, on tcgen05 (3xTF32,
 *     operands transposed on chip: x planes in tensor memory, gy planes swizzled in shared memory), split over M with a
 *     deterministic second-pass reduction.  gaux [(1 + n_aux), N_out] (optional): row 0 = column sums of gy (the bias
 *     gradient), row 1 + i = sum_m aux[m][i] * gy[m][:] (gradient of a rank-n_aux side term, n_aux <= 3).
 *     accumulate != 0 adds to gW / gaux.  N_out and K_in are multiples of 128; ldg, ldx multiples of 4.
 * ---------------------------------------------------------------------------------------------- */
size_t evrp_wgrad_workspace_bytes(void);
int evrp_wgrad_3xtf32(const float* gy, int ldg, int N_out, const float* x, int ldx, int K_in, int M, const float* aux,
                      int n_aux, float* gW, float* gaux, int accumulate, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * (9) train-mode BatchNorm1d(128) of the encoder (nets/GraphEncoder.py:142-150, batch statistics + running-stat update)
 *     over x [M,128], two phases so that ONE all-reduce between them turns the statistics into those of the global
 *     batch when instances are sharded over ranks:
 *       evrp_bn_stats            sums [257] fp64 = sum x | sum x^2 | M        (zeroed by the call)
 *       evrp_bn_apply            y = (x - mean) * invstd * weight + bias from the (all-reduced) sums; save [257] fp32 =
 *                                mean | invstd | n for the backward; running_mean / running_var (nullable) updated with
 *                                `momentum` and the unbiased variance like nn.BatchNorm1d
 *       evrp_bn_backward_reduce  sums [256] fp64 = sum g | sum g * xhat       (zeroed by the call; the local values are
 *                                the weight / bias gradients of this rank)
 *       evrp_bn_backward_apply   gx = weight * invstd * (g - sum_g / n - xhat * sum_gxhat / n) from the (all-reduced) sums
 * ---------------------------------------------------------------------------------------------- */
int evrp_bn_stats(const float* x, int M, double* sums, void* stream);
int evrp_bn_apply(const float* x, int M, const double* sums, const float* weight, const float* bias, float eps,
                  float momentum, float* running_mean, float* running_var, float* save, float* y, void* stream);
int evrp_bn_backward_reduce(const float* x, const float* g, int M, const float* save, double* sums, void* stream);
int evrp_bn_backward_apply(const float* x, const float* g, int M, const float* save, const float* weight,
                           const double* sums, float* gx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (10) step-query assembly of the teacher-forced decoder for all T recorded steps (nets/DRLModel.py:203-237,394):
 *      run_max[b,t,:] = max over the embeddings picked at steps < t (zeros at t = 0), base[b,t,:] = fixed_ctx[b,:] +
 *      emb[b, now_t, :] with now_0 = 0 (depot), now_t = pi[b,t-1]; and the backward: grad_emb [B,S,128] (written) and
 *      grad_fixed [B,128] from grad_run_max (nullable) and grad_base.  evrp_relu_mask: out = g * (y > 0), n % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
int evrp_route_context_forward(const float* emb, const float* fixed_ctx, const int64_t* pi, int B, int T, int S,
                               float* run_max, float* base, void* stream);
int evrp_route_context_backward(const float* emb, const int64_t* pi, const float* grad_run_max, const float* grad_base,
                                int B, int T, int S, float* grad_emb, float* grad_fixed, void* stream);
int evrp_relu_mask(const float* g, const float* y, float* out, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (11) init embedding on its own and its backward (training path): nets/DRLModel.py:192-200 `_init_embed` with the three
 *      parameters in nn.Linear layout (depot / station [128,2] + [128], customer [128,3] + [128]); the backward produces the
 *      six parameter gradients from grad_h [B,S,128] (10 weighted column reductions, deterministic two-pass sum).
 * ---------------------------------------------------------------------------------------------- */
int evrp_init_embed_forward(const float* static_xy, const float* dynamic, int B, int S, int F, const float* depot_w,
                            const float* depot_b, const float* station_w, const float* station_b, const float* cust_w,
                            const float* cust_b, float* h, void* stream);
size_t evrp_init_embed_backward_workspace_bytes(void);
int evrp_init_embed_backward(const float* static_xy, const float* dynamic, const float* grad_h, int B, int S, int F,
                             float* g_depot_w, float* g_depot_b, float* g_station_w, float* g_station_b, float* g_cust_w,
                             float* g_cust_b, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
