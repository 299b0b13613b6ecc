// Internal descriptor of the tcgen05 / TMEM / TMA bf16 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>

// Implicit-GEMM convolution: the A operand is gathered by TMA straight from an NHWC bf16 activation
// [img][h][w-slot][64-element slot] (rank-4 map, zero fill outside the image = padding).  Output rows are
// (img, oh, ow) with ow fastest; a 128-row tile is box_h output rows x 64 columns of box_img images.
//   k-block kb -> kh = kb / per_kh, j = kb % per_kh ; TMA coords (c, w, h, img) =
//   ((j % cch) * 64, j / cch + w_off, oh0 * sh - pad_h + kh, img0)
// box_w (0 = 64) output columns per output row: 64 or 128 (a power of two); 128 rows = box_w x box_h x box_img.
struct MrnbTcConv {
  int enabled;
  long dims[4];        // elements: {inner (>= 64), w slots, H, images}
  long strides[3];     // elements: {w slot, h, image}
  int box_h, box_img;  // 128 rows = 64 w-slots x box_h x box_img
  int sh;              // stride along h (TMA element stride)
  int rows_per_img, per_kh, cch, w_off;
  int imgs_per_group;
  int box_w, pad_h;
};

// Fused LSTM-cell epilogue (crnn.cu): the GEMM is gates = h W_hh^T for group g = (expert, direction) with the gate axis
// INTERLEAVED (column 4*j + {i,f,g,o}), so every float4 of the epilogue holds the four gates of one hidden unit.
// The epilogue adds the input pre-activation, applies the cell and writes c (fp32), h (bf16, next step's operand) and
// the [fwd | bwd] output row; nothing is written to `out`.  Needs M % 128 == 0.
struct MrnbTcLstm {
  int enabled;
  const void* pre; long pre_row, pre_e, pre_off[2];     // bf16 [expert][sample*64 + t][2 * 1024]: element strides / per-direction offsets
  float* cst; void* hst;                                  // [group][sample][256]
  void* rec; long rec_row, rec_e, rec_off[2];           // bf16 [expert][sample*64 + t][512]
  int B;
};

// Ragged classifier heads of all experts in ONE launch (modules/model.py:164,181: fc = Linear(256, C_i), C_i differs per
// expert): flat tile list over (expert, m tile, n tile).  The experts' bf16 weights are one stacked [sum C_i, K] matrix
// (expert e at row woff[e]); outputs / biases are per expert.  route != NULL: hard-routed inference
// (modules/model.py:383-393) -- a 128-row tile of expert e is computed only if the route sends one of its samples to e.
struct MrnbTcHeads {
  int n_experts;
  int tile_prefix[9];      // first flat tile of expert e; [n_experts] = total
  int n_tiles[8];          // n tiles (128 columns) of expert e
  int N[8];                // C_e
  int woff[8];             // first row of expert e in the stacked weight matrix
  float* out[8]; long ldo[8]; const float* bias[8];
  const int* route;        // device [n_samples] expert index per sample, or NULL
  int rows_per_sample, n_samples;
};

struct MrnbTcGemm {
  // out[g, m, n] = epi( sum_k A[g, m, k] * W[g, n, k] + bias[g, n] )     A, W: bf16, k-contiguous
  const void* A; long lda; long a_gstride;     // elements
  const void* W; long ldw; long w_gstride;
  const float* bias; long bias_gstride;
  void* out; long ldo; long o_gstride; int out_f32;    // fp32 or bf16 output
  const float* res;                                    // fp32 residual at the output address (out_f32 only)
  const float* rowscale; int rows_per_scale; long rowscale_gstride;   // DropPath: * rowscale[g*gs + m / rows_per_scale]
  int M, N, K, groups, gelu;
  int relu;                                            // ReLU on the biased accumulator (VGG convolutions)
  int pool4;                                           // bf16 out [M/4, N] = ReLU(max over each group of 4 consecutive rows + bias): 2x2 max-pool of window-major rows
  // optional fused LayerNorm of the fp32 output rows (N == 64 or 128): bf16 ln_out[g][m][N] = LN(out row) * gamma[g] + beta[g]
  void* ln_out; long ln_gstride; const float* ln_gamma; const float* ln_beta; float ln_eps;
  MrnbTcLstm lstm;     // optional fused LSTM cell (replaces the store)
  MrnbTcConv conv;     // optional: A is an implicit im2col view (A / lda / a_gstride ignored except A as base pointer)
};

int mrnb_tc_gemm(const MrnbTcGemm& g, cudaStream_t st);
// out[e][m, :C_e] = A[e][m, :] . Wall[woff[e] + n, :]^T + bias[e]   for every expert e (fp32 out); H.tile_prefix / n_tiles are filled here
int mrnb_tc_heads(const void* A, long lda, long a_gstride, const void* Wall, long w_rows, int M, int K, MrnbTcHeads H, cudaStream_t st);
// Persistent LSTM recurrence (all steps s_first .. CT-1 of one BidirectionalLSTM layer, every chain) in one cluster launch;
// MRNB_ERR_UNSUPPORTED = shape does not fit, fall back to one grouped GEMM per step.
int mrnb_tc_lstm_seq(const void* Whh, const void* pre, long pre_row, long pre_e, void* rec, long rec_row, long rec_e, float* cst,
                     void* h0, void* h1, int groups, int B, int CT, int s_first, cudaStream_t st);
