// Implicit-GEMM convolution / batched GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[128 x BLOCK_N] (fp32, TMEM)  =  sum_k  A[128 x 64] (bf16, smem via TMA)  *  B[BLOCK_N x 64]^T (bf16, smem via TMA)
//
// A rows are *output pixels*.  A is never materialised as an im2col matrix: the activation tensor is
// NHWC bf16 and described by a rank-4 TMA tensor map (c, w, h, n).  An output tile is a box of
// bn x bh x bw = 128 pixels; for filter tap (r, s) the producer loads the same box shifted by
// (r - pad, s - pad) and TMA zero-fills whatever falls outside the image, which *is* the conv padding.
// The K loop runs over "segments" (source tensor, taps, channel chunks) so that
//   * conv(cat[a, b])            = two segments over two tensor maps (no concat copy),
//   * conv3x3(h) + conv1x1(x)    = one accumulator (ResnetBlock conv2 + nin_shortcut, unet_small.py:128-136),
//   * stride-2 convs             = a tensor map with elementStrides = 2 (unet_small.py:69-73),
//   * plain / batched GEMMs      = one segment with a single tap (attention, 1x1 convs).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace dxmi {


// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is PER DEVICE: a "configured once per process" flag breaks a network living on
// cuda:1 while cuda:0 was configured first.  Usage: static DevFlags configured; if (!configured.test_and_set()) { set attribute }
struct DevFlags {
    bool done[64] = {};
    bool test() const {
        int d = 0;
        cudaGetDevice(&d);
        return done[d & 63];
    }
    void set() {
        int d = 0;
        cudaGetDevice(&d);
        done[d & 63] = true;
    }
};

// option "pdl": launch the hot kernels with programmatic stream serialization (see ptx.cuh); read at launch time
int pdl_enabled();
void set_pdl(int v);


struct GemmSeg {
    int map;      // index into a_map[]
    int ntaps;    // 1, 9 or 4 (up2 mode: 2x2)
    int nchunks;  // channels / 64
    int pad;      // 1 for 3x3 "same", 0 otherwise
};

enum GemmAct { ACT_NONE = 0, ACT_LRELU02 = 1, ACT_SILU = 2 };

struct ConvGemmParams {
    CUtensorMap a_map[3];
    CUtensorMap b_map;
    GemmSeg seg[3];
    int nseg;
    int stride;                // conv stride in w/h (1 or 2)
    int bw, bh, bn;            // output-pixel box of one tile: bw*bh*bn == 128
    int tiles_w, tiles_h;      // tiles per image along w / h
    int M_total;               // valid rows per batch entry
    int N_total;               // valid output columns
    int a_batched, b_batched;  // does blockIdx.z index the n coordinate of A / the batch coordinate of B?
    // ---- epilogue
    void* out;                 // bf16 (or fp32 if out_fp32) [row * ldo + col]
    long long out_batch_stride;
    int ldo;
    int out_fp32;
    int out_nchw;              // fp32 [image][col][pixel] (network output layout); one-tile-per-CTA kernel only
    const float* bias;         // per column (or per row if bias_along_m), may be null
    int bias_along_m;
    const float* rowvec;       // per-image per-column add, [image * ldrv + col], may be null
    int ldrv;
    int rows_per_image;
    const __nv_bfloat16* residual;  // same indexing as out (ld = ldr), may be null
    const __nv_bfloat16* gate;      // optional [rows, ldg]: out *= (gate > 0 ? 1 : 0.2)  (leaky-relu backward; generic epilogue)
    int ldg;
    long long res_batch_stride;
    int ldr;
    int act;
    float alpha;               // accumulator scale (applied first)
    int softmax;               // row softmax over the BLOCK_N columns (needs N_total == BLOCK_N)
    // ---- persistent kernel (gemm_tc2.cu) only
    CUtensorMap out_map;       // (cols, rows, batch) over `out`, box (128 bytes of columns, 128 rows, 1), SWIZZLE_128B
    CUtensorMap res_map;       // same geometry over `residual` (bf16)
    int m_tiles, n_tiles, batch_count;
    // halo mode (3x3 stride-1 convs on 32- / 64-wide maps): one TMA load of a [(rows+2) x (W+2) pixels x 64 channels] halo
    // tile per channel chunk serves all 9 taps through shifted smem descriptors; output rows are positions of the
    // zero-padded (W+2)-wide grid of one image (m_tiles = images * halo_tpi)
    int halo;                  // 0 = off
    int halo_W, halo_H;        // image width / height
    int halo_rows;             // rows of the TMA box (tile row span + 2)
    int halo_tpi;              // tiles per image = ceil(H * (W+2) / 128)
    int halo_a_stage;          // bytes of one A (halo) stage, multiple of 1024
    int halo_sb;               // B ring depth in halo mode
    // shift-3 mode (pair kernel, BLOCK_N = 128, 3x3 stride-1 convs whose 128-pixel tile is bh full image rows): per channel
    // chunk, THREE loads of a [(bh+2) rows x W x 64 ch] box - one per column shift s = -1, 0, +1 - serve the nine taps: tap
    // (r, s) is rows r*W .. r*W+127 of box s, a 1024-byte-aligned offset, so plain descriptors apply.  A bytes per tile and
    // chunk drop from 9 x 16 KB to 3 x (bh+2)*W*128 B: the SM's 64 B/clk ingest port, not the tensor pipe, bounds these layers.
    int shift3;                // 0 = off
    int s3_a_bytes;            // bytes of one A box = (bh+2) * W * 128
    int s3_row_bytes;          // W * 128: smem offset between vertically adjacent taps
    int s3_stages;             // ring depth in this mode
    int s3_m2;                 // 1: each CTA owns TWO vertically adjacent 128-row tiles per step (one A box of 2*bh+2 rows, the B tiles shared)
    CUtensorMap s3_map[3];     // per source: (c, w, h, n) map with box (64, W, bh+2, 1)
    // up2 mode (nearest-2x upsample folded into the following 3x3 convolution, unet_small.py:52-64, cm/unet.py:103-118): the four output
    // phases (py, px) = (Y & 1, X & 1) are four 2x2 convolutions of the LOW-resolution input with pre-summed weights (4/9 of the
    // FLOPs, no upsampled tensor).  The launch's batch index is the phase: B operand batched, A taps (r, q) in {0,1}^2 read the box
    // shifted by (r - 1 + py, q - 1 + px), tile row r = (n, y, x) is stored at output pixel (n, 2y + py, 2x + px).
    int up2;                   // 0 = off
    int up2_wmask;             // W - 1 (W = low-resolution width, a power of two)
    int up2_w2;                // 2 * W: output-row offset of phase row py
    int up2_spi;               // GroupNorm partial segments per low-resolution image (rows_per_image / stats_seg)
    int lean_epi;              // 1: bias-only tiles without statistics are drained straight from TMEM (gemm_epi.cuh epi_tile_lean)
    float* stats;              // GroupNorm partial sums of the bf16 outputs: [M_total/stats_seg][N_total][2] (sum, sumsq) or null
    int stats_seg;             // rows per partial: 32, 64 or 128 (a segment never straddles two images)
    long long* dbg_times;      // profiling only: per-CTA phase timestamps (globaltimer ns), 8 slots per CTA, or null
    int dbg_mode;              // profiling only: 1 = skip the MMAs (TMA ring throughput), 2 = skip the epilogue stores
};

// Launches the kernel; block_n in {32, 64, 128, 256}. Returns cudaError_t as int.
int launch_conv_gemm(const ConvGemmParams& p, int block_n, int m_tiles, int n_tiles, int batch, cudaStream_t stream);

// Host helper: encode a rank-4 (c, w, h, n) bf16 activation map. Strides in elements. Box = (64, bw*stride, bh*stride, bn).
int make_act_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, long long w_stride, long long h_stride,
                 long long n_stride, int bw, int bh, int bn, int stride);
// Host helper: encode a rank-3 (k, rows, batch) bf16 K-major operand map. Box = (64, box_rows, 1).
int make_mat_map(CUtensorMap* out, const void* base, int K, int rows, int batch, long long row_stride,
                 long long batch_stride, int box_rows);

// Persistent variant (gemm_tc2.cu): TMEM double buffering, TMA-store epilogue, fused GroupNorm partial statistics.
int launch_conv_gemm_v2(const ConvGemmParams& p, int block_n, cudaStream_t stream);
bool conv_gemm_v2_supported(const ConvGemmParams& p, int block_n);
// cta_group::2 pair variant (gemm_tc2p.cu); the B tensor map's box must hold block_n / 2 rows
int launch_conv_gemm_pair(const ConvGemmParams& p, int block_n, cudaStream_t stream);
void set_pair_resident_b(int v);
int pair_resident_b_enabled();  // 0 disables the weights-stationary variant of the pair kernel
int conv_gemm_v2_ring_bytes(int block_n);
int conv_gemm_pair_ring_bytes(int block_n);
// Host helper: (cols, rows, batch) map with a 128-byte x 128-row box for the epilogue (elem_bytes 2 = bf16, 4 = fp32).
int make_out_map(CUtensorMap* out, const void* base, int elem_bytes, int cols, int rows, int batch, long long row_stride,
                 long long batch_stride);

const char* gemm_last_error();

}  // namespace dxmi
