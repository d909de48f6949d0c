// Compile-time layer tables of the three reference models and the shared-memory /
// tensor-memory layout the fused network kernel derives from them.
//
// Layer tables restate model_utils/model.py:6-29 (V1), :32-61 (V2), :64-96 (V3) of the
// reference; every row is one conv_bn_relu call (model_utils/module.py:11-34) with BN
// already folded into kernel and bias by the host (model_utils/fold.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RCED_HD __host__ __device__
#else
#define RCED_HD
#endif

namespace rced {

constexpr int kBins = 129;        // frequency bins per frame
constexpr int kRS = 136;          // shared-memory row stride (floats): [8 halo | 128 bins], bin 128
                                  // lands on offset 0 of the next row, offsets 1..7 stay zero
constexpr int kRowBin0 = 8;       // offset of bin 0 inside a row
constexpr int kWS = 196;          // row stride of the "wide" layout read by the (1,129) layer
constexpr int kWideBin0 = 64;     // 64-float zero halo on each side (SAME pad of k=129)
constexpr int kFinalKP = 132;     // 129 taps padded to a multiple of 4
constexpr int kFramesPerCta = 4;  // one frame pipeline per SM sub-partition ...
constexpr int kSplit = 2;         // ... run by kSplit warps that share the frame's output channels
constexpr int kWarpsPerCta = kFramesPerCta * kSplit;   // warp w: frame slot w % 4, channel part w / 4
constexpr int kCombine = 136;     // per-slot scratch the two parts of the (1,129) layer meet in
constexpr int kMaxLayers = 16;

struct LSpec {
    int kh, kw, cin, cout;
    int save;    // skip slot this layer's output is parked in (-1: none)
    int add;     // skip slot added to this layer's output (-1: none)
    int relu;    // ReLU after (conv+bias [+skip])
    int after;   // 1: skip is added AFTER the ReLU and not rectified again (V3, model.py:75-76)
};

RCED_HD constexpr int num_layers(int arch) { return arch == 1 ? 10 : 16; }

RCED_HD constexpr LSpec spec(int arch, int i) {
    if (arch == 2) {   // FullyCNNSEModelV2, model.py:37-55
        switch (i) {
            case 0:  return {8, 11, 1, 10, 0, -1, 1, 0};
            case 1:  return {1, 7, 10, 12, 1, -1, 1, 0};
            case 2:  return {1, 5, 12, 14, 2, -1, 1, 0};
            case 3:  return {1, 5, 14, 15, 3, -1, 1, 0};
            case 4:  return {1, 5, 15, 19, 4, -1, 1, 0};
            case 5:  return {1, 5, 19, 21, 5, -1, 1, 0};
            case 6:  return {1, 7, 21, 23, 6, -1, 1, 0};
            case 7:  return {1, 11, 23, 25, -1, -1, 1, 0};
            case 8:  return {1, 7, 25, 23, -1, 6, 1, 0};
            case 9:  return {1, 5, 23, 21, -1, 5, 1, 0};
            case 10: return {1, 5, 21, 19, -1, 4, 1, 0};
            case 11: return {1, 5, 19, 15, -1, 3, 1, 0};
            case 12: return {1, 5, 15, 14, -1, 2, 1, 0};
            case 13: return {1, 7, 14, 12, -1, 1, 1, 0};
            case 14: return {1, 11, 12, 10, -1, 0, 1, 0};
            default: return {1, 129, 10, 1, -1, -1, 0, 0};
        }
    }
    if (arch == 3) {   // FullyCNNSEModelV3, model.py:68-91: five simple_RCED blocks + decode_final
        if (i == 15) return {1, 129, 8, 1, -1, -1, 0, 0};
        const int blk = i / 3, pos = i % 3;
        if (pos == 0) return {blk == 0 ? 8 : 1, 9, blk == 0 ? 1 : 8, 18, -1, -1, 1, 0};
        if (pos == 1) return {1, 5, 18, 30, -1, -1, 1, 0};
        // block output: CE1 -> slot 0, CE2 -> slot 1; CD1 adds CE2 (slot 1), CD2 adds CE1 (slot 0)
        return {1, 9, 30, 8, blk == 0 ? 0 : (blk == 1 ? 1 : -1), blk == 3 ? 1 : (blk == 4 ? 0 : -1), 1, 1};
    }
    // FullyCNNSEModel (V1), model.py:11-23
    switch (i) {
        case 0: return {8, 13, 1, 12, 0, -1, 1, 0};
        case 1: return {1, 11, 12, 16, 1, -1, 1, 0};
        case 2: return {1, 9, 16, 20, 2, -1, 1, 0};
        case 3: return {1, 7, 20, 24, 3, -1, 1, 0};
        case 4: return {1, 7, 24, 32, -1, -1, 1, 0};
        case 5: return {1, 7, 32, 24, -1, 3, 1, 0};
        case 6: return {1, 9, 24, 20, -1, 2, 1, 0};
        case 7: return {1, 11, 20, 16, -1, 1, 1, 0};
        case 8: return {1, 13, 16, 12, -1, 0, 1, 0};
        default: return {1, 129, 12, 1, -1, -1, 0, 0};
    }
}

RCED_HD constexpr int pad4(int x) { return (x + 3) & ~3; }
RCED_HD constexpr int pad2(int x) { return (x + 1) & ~1; }

// "input channels" the kernel iterates over: the 8 time taps for the first layer (cin == 1),
// the real channel count elsewhere.
RCED_HD constexpr int cin_eff(int arch, int i) { return i == 0 ? spec(arch, 0).kh : spec(arch, i).cin; }

// ---- canonical folded-weight order (what rced_create receives) --------------------------
RCED_HD constexpr int64_t folded_layer_floats(int arch, int i) {
    return (int64_t)spec(arch, i).kh * spec(arch, i).kw * spec(arch, i).cin * spec(arch, i).cout + spec(arch, i).cout;
}
RCED_HD constexpr int64_t folded_off(int arch, int i) {
    int64_t o = 0;
    for (int j = 0; j < i; ++j) o += folded_layer_floats(arch, j);
    return o;
}
RCED_HD constexpr int64_t folded_count(int arch) { return folded_off(arch, num_layers(arch)); }

// ---- channel split of a conv layer over the kSplit warps of a frame ---------------------
// part h owns output channels [h*ch_part, (h+1)*ch_part) (those >= cout are zero padding); the
// count is even because the inner loop works on channel PAIRS (packed fma.rn.f32x2).
RCED_HD constexpr int ch_part(int arch, int i) { return pad2((spec(arch, i).cout + kSplit - 1) / kSplit); }
// floats of one part's weights for one input channel: [kw][ch_part], padded to 16 bytes
RCED_HD constexpr int ci_block(int arch, int i) { return pad4(spec(arch, i).kw * ch_part(arch, i)); }

// ---- packed shared-memory weight image -----------------------------------------------
// conv layer i : W[part][cin_eff][ci_block] then bias[part][pad4(ch_part)]
// final layer  : W[cin][kFinalKP], S[cin][kFinalKP] (S[t] = W[t+1], the odd-bin pairing) then bias[4]
RCED_HD constexpr int packed_w_floats(int arch, int i) {
    return i == num_layers(arch) - 1 ? 2 * spec(arch, i).cin * kFinalKP
                                     : kSplit * cin_eff(arch, i) * ci_block(arch, i);
}
RCED_HD constexpr int packed_b_floats(int arch, int i) {
    return i == num_layers(arch) - 1 ? 4 : kSplit * pad4(ch_part(arch, i));
}
RCED_HD constexpr int packed_w_off(int arch, int i) {
    int o = 0;
    for (int j = 0; j < i; ++j) o += packed_w_floats(arch, j) + packed_b_floats(arch, j);
    return o;
}
RCED_HD constexpr int packed_b_off(int arch, int i) { return packed_w_off(arch, i) + packed_w_floats(arch, i); }
RCED_HD constexpr int packed_count(int arch) { return packed_w_off(arch, num_layers(arch)); }

// ---- per-warp activation slot --------------------------------------------------------
RCED_HD constexpr int max_channels(int arch) {
    int m = 0;
    for (int i = 0; i < num_layers(arch); ++i) {
        if (spec(arch, i).cout > m) m = spec(arch, i).cout;
        if (cin_eff(arch, i) > m) m = cin_eff(arch, i);
    }
    return m;
}
// floats of the wide layout the final layer reads
RCED_HD constexpr int wide_floats(int arch) { return spec(arch, num_layers(arch) - 1).cin * kWS + kWideBin0; }
// first row of the 8-row staging area the next frame's input is prefetched into while the final
// layer runs (must not overlap the wide layout)
RCED_HD constexpr int stage_row(int arch) { return (wide_floats(arch) + kRS - 1) / kRS; }
RCED_HD constexpr int slot_rows(int arch) {
    return max_channels(arch) > stage_row(arch) + 9 ? max_channels(arch) : stage_row(arch) + 9;
}
RCED_HD constexpr int slot_floats(int arch) { return slot_rows(arch) * kRS + 8 + kCombine; }
RCED_HD constexpr int combine_off(int arch) { return slot_rows(arch) * kRS + 8; }

// ---- tensor-memory columns of the skip slots ------------------------------------------------
// per part: 4 bins x ch_part channels + 1 value of bin 128 (lane == channel), parts back to back
RCED_HD constexpr int skip_part_cols(int arch, int slot) {
    for (int i = 0; i < num_layers(arch); ++i)
        if (spec(arch, i).save == slot) return 4 * ch_part(arch, i) + 1;
    return 0;
}
RCED_HD constexpr int skip_cols(int arch, int slot) { return kSplit * skip_part_cols(arch, slot); }
// a layer that adds skip slot s must split its channels exactly like the layer that saved it
RCED_HD constexpr bool skip_split_consistent(int arch) {
    for (int i = 0; i < num_layers(arch); ++i) {
        if (spec(arch, i).add < 0) continue;
        if (4 * ch_part(arch, i) + 1 != skip_part_cols(arch, spec(arch, i).add)) return false;
    }
    return true;
}
RCED_HD constexpr int skip_col_base(int arch, int slot) {
    int o = 0;
    for (int s = 0; s < slot; ++s) o += skip_cols(arch, s);
    return o;
}
RCED_HD constexpr int skip_total_cols(int arch) { return skip_col_base(arch, 8); }

// ---- roofline numerators -------------------------------------------------------------
RCED_HD constexpr int64_t mac_per_frame(int arch, bool valid_only) {
    int64_t total = 0;
    for (int i = 0; i < num_layers(arch); ++i) {
        const LSpec s = spec(arch, i);
        const int pl = (s.kw - 1) / 2;
        int64_t taps = 0;
        if (valid_only) {
            for (int f = 0; f < kBins; ++f)
                for (int k = 0; k < s.kw; ++k)
                    if (f + k - pl >= 0 && f + k - pl < kBins) ++taps;
        } else {
            taps = (int64_t)kBins * s.kw;
        }
        total += taps * s.kh * s.cin * s.cout;
    }
    return total;
}

static_assert(skip_split_consistent(1) && skip_split_consistent(2) && skip_split_consistent(3), "skip split mismatch");
static_assert(skip_total_cols(1) <= 512 && skip_total_cols(2) <= 512 && skip_total_cols(3) <= 512,
              "skip tensors must fit the 512 tensor-memory columns of one lane quadrant");

}  // namespace rced
