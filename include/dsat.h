/*
 * dsat.h -- C ABI of the B200-native DiffusionSAT sampling hot path (libdsat.so).
 *
 * The reference (LUMII-Syslab/DiffusionSAT) has no FFI of its own: its boundary for this path is the
 * Python API  DiffusionSampler(model_path, dimacs).samples(n)  and  QuerySAT.diffusion_step(...).
 * Each entry point below names the reference code it replaces (file:line into the reference tree).
 * The Python mirror of that API (diffusionsat_b200/) binds these symbols with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 or a negative dsat_status and
 * leaves a message for dsat_last_error(); no exceptions cross the ABI; "host" buffers are caller
 * owned and copied, "dev" buffers are device pointers of the context's device; one context per GPU,
 * not thread-safe; all work is issued on the context's stream; there is no CPU fallback.
 *
 * Row order of every per-variable array is the reference's batch order: variable v of graph g of
 * chain c sits at  c*n_vars + v  with n_vars the unit's variable total (data/dimac.py:239-241,
 * data/SatSpecifics.py:22-35); graphs are numbered  c*n_graphs + g.
 */
#ifndef DSAT_H_
#define DSAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsat_ctx dsat_ctx;

enum dsat_status {
    DSAT_OK = 0,
    DSAT_ERR_ARG = -1,      /* bad argument / call order */
    DSAT_ERR_CUDA = -2,     /* CUDA runtime error, see dsat_last_error */
    DSAT_ERR_STATE = -3,    /* model or graph not set */
    DSAT_ERR_UNSUPPORTED = -4
};

/* Arithmetic of the twelve Dense layers (everything else is fp32 in every mode):
 *   DSAT_F32          fp32 FMA on the CUDA cores (bit-level cross-check of the tensor-core paths)
 *   DSAT_F32_TC       fp32-accurate on the tcgen05 tensor cores: every operand is carried as two bf16 planes (hi, lo) and
 *                     every product is three bf16 MMAs into one fp32 accumulator (~1e-5 relative); the default of the
 *                     Python drop-in classes; matches the reference's fp32 Dense layers within the 1e-3 tolerance
 *   DSAT_BF16         plain bf16 operands and bf16 activation storage, one tcgen05 kernel per MLP (stated separately)
 *   DSAT_BF16_UNFUSED the same with one tcgen05 kernel per Dense layer */
enum dsat_dtype { DSAT_F32 = 0, DSAT_BF16 = 1, DSAT_BF16_UNFUSED = 2, DSAT_F32_TC = 3 };

/* debug/parity access to the activation buffers of one round (dsat_debug_read / dsat_debug_write) */
enum dsat_buffer {
    DSAT_BUF_VROW = 0,    /* [N, F+16+3Q]  variables | aux16 | variables_grad | loss_pos | loss_neg */
    DSAT_BUF_CROW = 1,    /* [M, F+2Q]     clause_state | clause_messages | 4*clauses_loss          */
    DSAT_BUF_H1 = 2,      /* [N, Hq+4Q]    hidden of variables_query | first hidden of lit_query     */
    DSAT_BUF_H2 = 3,      /* [N, 4Q]       second hidden of lit_query                                 */
    DSAT_BUF_QS = 4,      /* [N, 3Q]       query | softplus(query) | softplus(-query)                 */
    DSAT_BUF_LIT = 5,     /* [N, 2Q]       lit_query output (positive | negative literal features)    */
    DSAT_BUF_CH = 6,      /* [M, Hc]       hidden of clause_update                                    */
    DSAT_BUF_COUT = 7,    /* [M, Q+F]      clause_update output (message to literals | new value)     */
    DSAT_BUF_U1 = 8,      /* [N, Hu] */
    DSAT_BUF_U2 = 9,      /* [N, Hu] */
    DSAT_BUF_UOUT = 10,   /* [N, F]        update_gate output before PairNorm                          */
    DSAT_BUF_SPRE = 11,   /* [N, F]        variables after PairNorm+residual, before the 0.2/0.8 carry  */
    DSAT_BUF_O1 = 12,     /* [N, Ho] */
    DSAT_BUF_LOGITS = 13, /* [N, 16]       8 logit maps + padding                                       */
    DSAT_BUF_OUT = 14,    /* [N]           selected logit per variable (out_logits)                     */
    DSAT_BUF_X = 15,      /* [N, 2]        diffusion state x                                            */
    DSAT_BUF_COUNT = 16
};

int dsat_version(void);

/* Lifetime.  One context per GPU.  Replaces the implicit TF runtime/device placement of
 * satuniformity/DiffusionSampler.py:197-213. */
int dsat_create(int device, dsat_ctx** out);
void dsat_destroy(dsat_ctx* ctx);
const char* dsat_last_error(const dsat_ctx* ctx);

/* Use a caller stream (cudaStream_t passed as void*) for all work; NULL restores the own stream. */
int dsat_set_stream(dsat_ctx* ctx, void* cuda_stream);
int dsat_synchronize(dsat_ctx* ctx);
/* CUDA-event timing on the context's stream (bench.py): begin records, end records+syncs, ms out. */
int dsat_timer_begin(dsat_ctx* ctx);
int dsat_timer_end(dsat_ctx* ctx, float* elapsed_ms);
/* kernels launched by this context since creation (bench.py "gpu_launches") */
long long dsat_launch_count(const dsat_ctx* ctx);

/* Weights of the twelve Dense layers in the order variables_query/0,1  lit_query/0,1,2
 * clause_update/0,1  update_gate/0,1,2  variables_output/0,1; kernels[i] is row-major [in,out] fp32,
 * biases[i] is [out].  Replaces the checkpoint restore of DiffusionSampler.py:215-227 and the layer
 * construction of model/query_sat.py:117-122 (model/mlp.py:23-24,39). */
int dsat_set_model(dsat_ctx* ctx, int n_layers, const float* const* kernels, const float* const* biases,
                   const int* in_dims, const int* out_dims);

/* dtype of the MLP path, see enum dsat_dtype */
int dsat_set_precision(dsat_ctx* ctx, int dtype);
/* The active dtype.  A new context starts in DSAT_F32_TC; dsat_set_model switches it to DSAT_F32 (CUDA cores, the same
 * fp32 results) for layer widths the split-precision kernels do not tile (feature_maps or query_maps = 256). */
int dsat_get_precision(const dsat_ctx* ctx);

/* Unit graph shared by all chains.  cl_lit holds literal codes 2*var+sign (var 0-based in the unit);
 * lit_rowptr is indexed by literal code.  var_seg/clause_seg [n_graphs+1] delimit the formulas of a
 * disjoint union.  group_graphs = graphs per early-exit batch (reference: all graphs of one TF batch,
 * model/query_sat.py:330-338); 0 means "all graphs of this context".
 * Replaces data/dimac.py:14-18,213-260 + data/SatSpecifics.py:21-69 (adjacency construction) and the
 * degree weights of model/query_sat.py:193-197. */
int dsat_set_graph(dsat_ctx* ctx, int n_vars, int n_clauses, int nnz,
                   const int32_t* cl_rowptr, const int32_t* cl_lit,
                   const int32_t* lit_rowptr, const int32_t* lit_clause,
                   int n_graphs, const int32_t* var_seg, const int32_t* clause_seg,
                   int n_chains, int group_graphs);

/* One model call = QuerySAT.diffusion_step / call(training=False)  (model/query_sat.py:133-184,
 * 186-373, 467-481).  Host buffers: noisy_num [N,2]; labels [N] int32 or NULL (drawn from Philox,
 * the reference draws them with tf.random at :145); normals [rounds,N,4] or NULL (Philox; reference
 * tf.random.normal at :239); prediction_out [N] logits; steps_taken/loss [n_groups] (may be NULL). */
int dsat_model_call(dsat_ctx* ctx, float noise_scale, const float* noisy_num, const int32_t* labels,
                    const float* normals, int rounds, uint64_t seed, uint64_t chain_offset,
                    float* prediction_out, int32_t* steps_taken, float* loss);

/* Whole reverse-diffusion run for all chains of the context = diffusion() + the per-graph
 * int-encoding / SAT check of samples()  (satuniformity/DiffusionSampler.py:78-191, 283-303;
 * utils/VariableAssignment.py:63-90).  Optional injected noise (host): uniforms [steps,N],
 * labels [steps,N] int32, normals [steps,rounds,N,4]; NULL = Philox keyed by (seed, global chain).
 * Outputs (host): packed [n_total_graphs, words] little-endian 64-bit words, x1 = bit 0;
 * is_sat [n_total_graphs]; latch_step [n_total_graphs] (-1 = never satisfied);
 * sat_any_step [n_total_graphs] (cum_accuracy flags of diffusion(), may be NULL). */
int dsat_sample(dsat_ctx* ctx, int n_steps, int n_rounds, uint64_t seed, uint64_t chain_offset,
                const float* uniforms, const int32_t* labels, const float* normals,
                uint64_t* packed, uint8_t* is_sat, int32_t* latch_step, uint8_t* sat_any_step);
/* The same run split for resident benchmarking: enqueue only (no host copies, no sync), then fetch. */
int dsat_sample_enqueue(dsat_ctx* ctx, int n_steps, int n_rounds, uint64_t seed, uint64_t chain_offset);
int dsat_sample_fetch(dsat_ctx* ctx, uint64_t* packed, uint8_t* is_sat, int32_t* latch_step,
                      uint8_t* sat_any_step);
int dsat_words_per_graph(const dsat_ctx* ctx);

/* Stand-alone segment-sum SpMM on device buffers (message-passing roofline sweeps):
 * direction 0: clause <- literal   Y[c,j,:] = rev_w[j]  * sum_{lit in j} X[c,lit,:]   X [chains,2n,feat]
 * direction 1: literal <- clause   Y[c,l,:] = deg_w[l] * sum_{j contains l} X[c,j,:]  X [chains,m,feat]
 * (tf.sparse.sparse_dense_matmul call sites model/query_sat.py:255,269).  feat in {64,128,256}. */
int dsat_spmm(dsat_ctx* ctx, int direction, const void* x_dev, void* y_dev, int feat, int dtype, int chains);

/* Per-kernel-class device time of `rounds` message-passing rounds, measured with CUDA events on the
 * context's stream (bench.py roofline).  Classes 0..10 are the eleven linear ops in launch order
 * (v1->hidden, query out, lit 2, lit 3, clause 1, clause 2, update 1, 2, 3, output 1, 2), then
 * clause gather, literal gather, clause PairNorm, variable PairNorm, head, noise.
 * class_ms / class_launches have dsat_profile_classes() entries. */
int dsat_profile_classes(void);
/* clock64 wait/work breakdown of CTA 0 of one whole-MLP kernel (0 query .. 4 output); 16 counters */
int dsat_profile_fused(dsat_ctx* ctx, int which, long long* counters16);
int dsat_profile_rounds(dsat_ctx* ctx, int rounds, uint64_t seed, float* class_ms, int32_t* class_launches);

/* Stand-alone run of the tcgen05 linear kernel on host data (parity of the tensor-core MLP path,
 * model/mlp.py:42-50): out = epi(bf16(a) @ bf16(w) + bias); a [rows,K], w [K,N], out [rows,N] fp32
 * ([rows,3N] for epi 2 = query epilogue with the softplus pair); epi 0 linear, 1 leaky-relu 0.2. */
int dsat_tc_linear_test(dsat_ctx* ctx, int rows, int K, int N, const float* a_host, const float* w_host,
                        const float* bias_host, int epi, int out_bf16, float* out_host);

/* Parity hooks: run the pieces of one model call separately and read/write activation buffers. */
int dsat_debug_begin(dsat_ctx* ctx, float noise_scale, const float* noisy_num, const int32_t* labels);
int dsat_debug_round(dsat_ctx* ctx, int round, const float* normals /* [N,4] host */);
int dsat_debug_dims(const dsat_ctx* ctx, int buffer, long long* rows, int* ld);
int dsat_debug_read(dsat_ctx* ctx, int buffer, float* host_out, long long count);
int dsat_debug_write(dsat_ctx* ctx, int buffer, const float* host_in, long long count);
int dsat_debug_groups(dsat_ctx* ctx, int32_t* done, int32_t* steps_taken, float* loss_sum,
                      int32_t* graph_sat, int32_t* graph_map);
/* One MLP alone in the active precision on the current contents of its input buffer (model/query_sat.py:117-122):
 * which = 0 variables_query (VROW -> QS), 1 lit_query (VROW -> LIT), 2 clause_update (CROW -> COUT),
 * 3 update_gate (VROW -> UOUT), 4 variables_output (SPRE -> LOGITS). */
int dsat_debug_mlp(dsat_ctx* ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* DSAT_H_ */
