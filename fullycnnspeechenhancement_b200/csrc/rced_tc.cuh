// Layout of the tensor-core variant of the fused network kernel (rced_net_tc.cu), shared by
// the kernel, the host-side weight packer and the tests.
//
// The conv layers of model_utils/model.py run as implicit GEMMs on tcgen05 WITHOUT an im2col
// copy: activations live in shared memory as FP16 "chunk planes" A[cg][row][8 channels]
// (16 bytes per row and channel group), rows = (frame, bin) flattened with a frame stride of
// 136 (129 bins + 7 zero rows, the SAME-padding halo of both neighbours), and every filter tap
// is one K-major, no-swizzle matrix descriptor whose start address is the plane shifted by
// (tap - pad) rows (tools/umma_probe.cu proves the addressing).  FP32 accuracy comes from an
// error-compensated split x = hi + lo (two FP16 numbers, 22 significant bits):
//   x*w ~= hi*Whi + hi*Wlo + lo*Whi     (FP32 accumulation in tensor memory)
// issued as two instructions per K step: A_hi x [Whi | Wlo] (N = 2*NP) and A_lo x Whi (N = NP).
//
// Range (the FP16 exponent has 5 bits).  Three power-of-two scalings, all exact in FP32, keep every
// FP16 operand in its normal range whatever the scale of the input or of the weights:
//  * the residuals are stored times 2^11 (lo' = 2^11 (x - hi), Wlo' = 2^11 (w - Whi)), so they have the
//    magnitude of the value itself instead of falling into the FP16 subnormals; the two cross products
//    hi*Wlo' + lo'*Whi accumulate in their own columns [NP, 2 NP) and the epilogue adds them times 2^-11;
//  * a step's weights are scaled by 2^kw when its largest weight lies outside [2^-7, 2^3) (conv steps: to
//    [2^-2, 2^-1); BN folds with extreme gamma or variance) and always for the output layer (to
//    [2^12, 2^13), kStepWeightTop); the domain of every later activation shifts by 2^kw with it -- the
//    host pre-scales the biases, the skip additions multiply by the power of two between the two
//    domains (table behind the biases), and the output is multiplied by the inverse of the total;
//  * every FRAME of a batch is computed in its own scaled domain x' = s x, s = 2^k chosen by the prefetch
//    warp so that max(|input of the frame|, largest |bias|) s lies in [2^3, 2^4) (kFrameTop): rows of
//    different frames never meet in a conv layer (the time taps of the first layer are the channels of
//    the frame's own rows), ReLU and the skip additions commute with s, biases are added times s and
//    the output is multiplied by 1/s.
// An activation keeps its 22 bits while it lies within [2^-14, 65504] of the scaled domain, i.e. between
// 4e-6 and 4000 times the frame's reference magnitude; beyond 65504 the range guard hands the call to
// the FP32 FFMA kernel.
//
// The (1,129) output layer runs "taps in N with row-shifted accumulation": tap j = 32 i + n
// (i = 0..4, n = 0..31), E[r][n] = sum_i X[r + 32 (i - 2)] . W[32 i + n] -- five MMAs whose A
// descriptors are shifted by 32 (i - 2) rows accumulate into the same 32 columns -- and
// out[b] = sum_n E[b + n][n] (a 32-tap diagonal sum in the epilogue instead of a 129-tap one).
// Rows shifted by up to 64 would reach the neighbouring frame (frame stride 136), therefore the
// last conv layer writes its output twice, as an "even frames only" copy (planes 0, 1) and an "odd
// frames only" copy (planes 2, 3, dead by then), and each parity gets its own 32 accumulator columns.
#pragma once
#include "rced_arch.cuh"

namespace rced {
namespace tc {

constexpr int kFB = 7;                        // frames per CTA batch
constexpr int kFS = 136;                      // rows per frame in the flattened row space
constexpr int kLead = 8;                      // zero rows in front of the first frame
constexpr int kTiles = 8;                     // M = 128 row tiles per batch
constexpr int kRows = kTiles * 128;           // rows covered by the tiles (7 * 136 = 952 used)
constexpr int kRowsAlloc = kLead + kRows + 8; // + zero rows behind the last tile
constexpr int kPlane16 = kRowsAlloc;          // plane stride in 16-byte units
constexpr int kPlanes = 4;                    // channel groups of 8 (<= 32 channels)
constexpr int kLo16 = kPlanes * kPlane16;     // offset of the lo planes behind the hi planes
constexpr int kActBytes = 2 * kPlanes * kPlane16 * 16;
constexpr int kAccCols = 64;                  // tensor-memory columns per tile
constexpr int kFinalN = 32;                   // taps per row-shifted block of the (1,129) layer (N of its instructions)
constexpr int kFinalShifts = 5;               // blocks: row shifts -64, -32, 0, 32, 64
constexpr int kFrontRows = 64;                // zero rows in front of plane 0 (the -64 shift of row tile 0 reads them)
constexpr int kFrontPad = kFrontRows * 16;    // ... in bytes
#ifndef RCED_TC_ISSUERS
#define RCED_TC_ISSUERS 3
#endif
constexpr int kIssuers = RCED_TC_ISSUERS;     // MMA-issuing threads (2 .. 4)
constexpr int kCtrlWarps = 3 + kIssuers;      // warps 0, 3, 5, 6: MMA issue (row tiles round robin), 1: weight producer, 2: dependency scout,
                                              // 4: prefetch of the next batch's utterance bounds and input rows
constexpr int kEpiWarps = 16;                 // groups of four warps (one per tensor-memory lane quadrant)
constexpr int kGroups = kEpiWarps / 4;        // group g takes the row tiles t = g, g + kGroups, ...
constexpr int kThreads = 32 * (kCtrlWarps + kEpiWarps);
constexpr int kTraceEvents = 8;               // clock stamps per (step, tile) of the development trace
constexpr int kInRows = kFB + 7;              // input rows a batch reads: frames g0-3 .. g0+kFB+3
constexpr int kInStride = 132;                // floats per prefetched input row
constexpr int kLoShift = 11;                  // residuals (activations and weights) are stored times 2^kLoShift
constexpr int kStepWeightTop = 13;            // the output layer's weights are scaled so that max |w| lies in [2^12, 2^13)
constexpr int kFrameTop = 4;                  // a frame's scale puts max(|input|, |bias|) into [2^3, 2^4)
constexpr int kScaleRing = 4;                 // batches whose frame scales are kept (the epilogue of a batch reads them until it ends)

static_assert(kFB * kFS <= kRows, "frames of a batch must fit the row tiles");
static_assert(kFS - kBins >= 6 && kLead >= 6, "zero rows must cover the widest SAME pad (kw = 13)");

// steps: conv layers 0 .. NL-2, then the (1,129) layer
RCED_HD constexpr int n_steps(int arch) { return num_layers(arch); }
RCED_HD constexpr bool is_final(int arch, int s) { return s >= num_layers(arch) - 1; }
RCED_HD constexpr int step_layer(int arch, int s) { return is_final(arch, s) ? num_layers(arch) - 1 : s; }
// NP: output channels of a conv step padded to a multiple of 8.  The first instruction of a unit (N = 2 NP, a multiple
// of 16 as M = 128 requires) writes [hi*Whi | hi*Wlo'] to columns [0, 2 NP); the second one has N = NP rounded up to 16
// and starts at column NP: when NP is 8 or 24 its last 8 columns hold products with the first rows of the Wlo' block --
// they land beyond column 2 NP, inside the tile's 64 columns, and are never read.  (Round 1 padded to 16 / 32: the B
// fetch of a 19..24-channel step cost 32 + 16 rows instead of 24 + 16.)
RCED_HD constexpr int step_np(int arch, int s) {
    return is_final(arch, s) ? kFinalN : (spec(arch, s).cout + 7) / 8 * 8;
}
RCED_HD constexpr int step_n2(int arch, int s) { return (step_np(arch, s) + 15) / 16 * 16; }   // N of the second instruction
static_assert(step_np(1, 4) + step_n2(1, 4) <= kAccCols && step_np(3, 1) + step_n2(3, 1) <= kAccCols, "accumulator columns per tile");
RCED_HD constexpr int step_groups(int arch, int s) {
    return is_final(arch, s) ? (spec(arch, num_layers(arch) - 1).cin + 7) / 8 : (cin_eff(arch, s) + 7) / 8;
}
RCED_HD constexpr int step_chunks(int arch, int s) {
    return is_final(arch, s) ? step_groups(arch, s) : spec(arch, s).kw * step_groups(arch, s);
}
// units: pairs of (tap, group) chunks of a conv layer; the row-shifted blocks of the output layer (whose
// two chunks are the channel groups, at most 16 channels)
RCED_HD constexpr int step_units(int arch, int s) { return is_final(arch, s) ? kFinalShifts : (step_chunks(arch, s) + 1) / 2; }
// B tile of one unit: [2 chunks][rows][8 halfs]; conv steps: rows = 2*NP (Whi | Wlo'), output layer: 3*NP
// (Whi | Wlo | Whi 2^-11: its three products share one accumulator block, so the scaled residual of the
// activations meets a copy of the weights that carries the 2^-11)
RCED_HD constexpr int step_tile_rows(int arch, int s) { return (is_final(arch, s) ? 3 : 2) * step_np(arch, s); }
RCED_HD constexpr int step_tile_bytes(int arch, int s) { return 2 * step_tile_rows(arch, s) * 16; }
RCED_HD constexpr int step_w_bytes(int arch, int s) { return step_units(arch, s) * step_tile_bytes(arch, s); }
RCED_HD constexpr int step_w_off(int arch, int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) o += step_w_bytes(arch, i);
    return o;
}
RCED_HD constexpr int w_image_bytes(int arch) { return step_w_off(arch, n_steps(arch)); }
RCED_HD constexpr int max_step_w_bytes(int arch) {
    int m = 0;
    for (int i = 0; i < n_steps(arch); ++i)
        if (step_w_bytes(arch, i) > m) m = step_w_bytes(arch, i);
    return m;
}
RCED_HD constexpr int unit_base(int arch, int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) o += step_units(arch, i);
    return o;
}
RCED_HD constexpr int total_units(int arch) { return unit_base(arch, n_steps(arch)); }
RCED_HD constexpr int max_units(int arch) {
    int m = 0;
    for (int i = 0; i < n_steps(arch); ++i)
        if (step_units(arch, i) > m) m = step_units(arch, i);
    return m;
}
constexpr int kMaxUnits = 18;   // most units of a step
constexpr int kTabStride = 20;  // words per step in the constant-memory A-descriptor table
static_assert(n_steps(1) % 2 == 0 && n_steps(2) % 2 == 0 && n_steps(3) % 2 == 0, "step s must always use weight buffer s & 1");
static_assert(max_units(1) <= kMaxUnits && max_units(2) <= kMaxUnits && max_units(3) <= kMaxUnits, "raise kMaxUnits");

// A-operand addressing of chunk c of conv step s, in 16-byte units relative to row (kLead + 128 t)
// of plane 0: channel group g = c / kw lives in plane g, tap j = c % kw reads rows shifted by j - pad
RCED_HD constexpr int chunk_off16(int arch, int s, int c) {
    const int kw = spec(arch, s).kw;
    return (c / kw) * kPlane16 + (c % kw) - (kw - 1) / 2;
}
// unit u of step s: offset of its first chunk and distance to its second one.  A conv unit without
// a second chunk points at the next row (its weights are zero); output-layer unit u is the block
// shifted by 32 (u - 2) rows, its chunks are the two channel groups (planes 0, 1 of the even-frame
// copy; the odd-frame copy is 2 planes further)
RCED_HD constexpr int unit_off16(int arch, int s, int u) {
    return is_final(arch, s) ? kFinalN * (u - kFinalShifts / 2) : chunk_off16(arch, s, 2 * u);
}
RCED_HD constexpr int unit_lbo16(int arch, int s, int u) {
    if (is_final(arch, s)) return kPlane16;
    return 2 * u + 1 < step_chunks(arch, s) ? chunk_off16(arch, s, 2 * u + 1) - chunk_off16(arch, s, 2 * u) : 1;
}
static_assert(spec(1, num_layers(1) - 1).cin <= 16 && spec(2, num_layers(2) - 1).cin <= 16 && spec(3, num_layers(3) - 1).cin <= 16,
              "the output layer's input must fit two channel groups");
static_assert(spec(1, num_layers(1) - 1).kw <= kFinalN * kFinalShifts - kFinalN + 1 && kFinalN * (kFinalShifts / 2) == 64,
              "tap j = 32 i + n must cover the 129 taps around the centre tap 64");

// ---- skip tensors: FP32 in a per-CTA global scratch, [group of 8 channels][half][row][4] ----
RCED_HD constexpr int skip_c8(int arch, int slot) {
    for (int i = 0; i < num_layers(arch); ++i)
        if (spec(arch, i).save == slot) return (spec(arch, i).cout + 7) / 8;
    return 0;
}
RCED_HD constexpr int skip_c8_base(int arch, int slot) {
    int o = 0;
    for (int s = 0; s < slot; ++s) o += skip_c8(arch, s);
    return o;
}
RCED_HD constexpr int skip_c8_total(int arch) { return skip_c8_base(arch, 8); }
RCED_HD constexpr size_t skip_floats_per_cta(int arch) { return (size_t)skip_c8_total(arch) * kRows * 8; }

// ---- shared memory carve-up (bytes) ---------------------------------------------------------
RCED_HD constexpr int pad128(int x) { return (x + 127) & ~127; }
constexpr int smem_act_off = kFrontPad;   // activation planes, behind the zero rows
RCED_HD constexpr int smem_w_off(int arch, int buf) { return kFrontPad + kActBytes + buf * pad128(max_step_w_bytes(arch)); }
// bias table (global and shared): float[n_steps + 1][32]; row s: biases of step s, in the step's domain; row
// n_steps: [s] = conv step s: power of two between the domain of the skip tensor it adds and its own, output
// layer: inverse of the total weight scale; [kBiasRefSlot] = largest |bias| of the model (in its domain)
constexpr int kBiasRefSlot = 16;
RCED_HD constexpr int bias_floats(int arch) { return 32 * (n_steps(arch) + 1); }
RCED_HD constexpr int smem_bias_off(int arch) { return smem_w_off(arch, 2); }                                   // float[n_steps + 1][32]
RCED_HD constexpr int smem_out_off(int arch) { return smem_bias_off(arch) + 4 * bias_floats(arch); }           // float[2][kRows]: the two partial sums of every output row
// 256 bytes: int[2][8] time-tap masks of the prefetched batches, float[kScaleRing][8] frame scales (+64),
// float[16] row maxima of the batch being prefetched (+192)
RCED_HD constexpr int smem_bnd_off(int arch) { return smem_out_off(arch) + 2 * kRows * 4; }
RCED_HD constexpr int smem_bar_off(int arch) { return smem_bnd_off(arch) + 256; }                              // mbarriers
RCED_HD constexpr int smem_epi_off(int arch) { return smem_bar_off(arch) + 256; }                             // EpiStep[n_steps]
RCED_HD constexpr int smem_in_off(int arch) { return smem_epi_off(arch) + pad128(32 * n_steps(arch)); }       // float[2][kInRows][kInStride]
RCED_HD constexpr int smem_total(int arch) { return smem_in_off(arch) + 2 * kInRows * kInStride * 4; }

// mbarrier slots (8 bytes each) inside the barrier block
constexpr int kBarAccFull = 0;     // [kTiles]  tcgen05.commit after the last MMA of (step, tile)
constexpr int kBarActReady = 8;    // [kTiles]  epilogue of (step, tile) done (planes written, accumulator free)
constexpr int kBarWFull = 16;      // [2]       weights of a step landed in buffer b
constexpr int kBarWFree = 18;      // [2]       MMAs reading buffer b complete
constexpr int kBarInReady = 20;    //           layer-0 input of the batch staged
constexpr int kBarFinalDone = 21;  //           every MMA of the batch's output layer complete (plane 0 may be overwritten)
constexpr int kNextInSlot = 25;    //           u32 count of batches whose bounds / input rows the producer has prefetched
constexpr int kIssuedSlot = 26;    //           u32 count of row tiles whose MMAs have all been issued (and committed)
constexpr int kFlagSlot = 24;      //           u32 progress counter published by the dependency scout

}  // namespace tc
}  // namespace rced
