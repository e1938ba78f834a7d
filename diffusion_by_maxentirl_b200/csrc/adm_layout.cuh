// Module structure of the ADM / EDM U-Net (models/cm/unet.py:600-737), shared by the inference plan (engine_adm.cu) and the
// training plan (engine_train_adm.cu).
#pragma once
#include <vector>

#include "engine.cuh"

namespace dxmi {

enum LayerKind { L_CONV, L_RES, L_DOWN, L_UP, L_ATTN };
struct Layer {
    LayerKind kind;
    int cin, cout;
};
typedef std::vector<std::vector<Layer>> BlockList;

inline bool attn_at(const dxmi_arch_desc& a, int ds) {
    for (int i = 0; i < a.n_attn; ++i)
        if (a.attn_resolutions[i] == ds) return true;
    return false;
}

// Module structure of UNetModel.__init__ (models/cm/unet.py:600-737).
inline void adm_layout(const dxmi_arch_desc& a, BlockList& inputs, BlockList& outputs, int* mid_ch) {
    const int mc = a.ch;
    int ch = a.ch_mult[0] * mc;
    inputs.push_back({{L_CONV, a.in_channels, ch}});
    std::vector<int> chans{ch};
    int ds = 1;
    for (int level = 0; level < a.n_levels; ++level) {
        for (int i = 0; i < a.num_res_blocks; ++i) {
            const int out = a.ch_mult[level] * mc;
            std::vector<Layer> layers{{L_RES, ch, out}};
            ch = out;
            if (attn_at(a, ds)) layers.push_back({L_ATTN, ch, ch});
            inputs.push_back(layers);
            chans.push_back(ch);
        }
        if (level != a.n_levels - 1) {
            inputs.push_back({{L_DOWN, ch, ch}});
            chans.push_back(ch);
            ds *= 2;
        }
    }
    *mid_ch = ch;
    for (int level = a.n_levels - 1; level >= 0; --level) {
        for (int i = 0; i <= a.num_res_blocks; ++i) {
            const int ich = chans.back();
            chans.pop_back();
            const int out = a.ch_mult[level] * mc;
            std::vector<Layer> layers{{L_RES, ch + ich, out}};
            ch = out;
            if (attn_at(a, ds)) layers.push_back({L_ATTN, ch, ch});
            if (level && i == a.num_res_blocks) {
                layers.push_back({L_UP, ch, ch});
                ds /= 2;
            }
            outputs.push_back(layers);
        }
    }
}

}  // namespace dxmi
