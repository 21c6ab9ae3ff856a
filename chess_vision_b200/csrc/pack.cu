// Host-side weight packer of the C-ABI: reference state_dict tensors -> the packed fp32 blob cv_square_load_weights consumes.
// Restates chess_vision_b200/weights.py (pack_state_dict) operation by operation in IEEE double arithmetic, so both packers emit
// the same bits (tests/test_pack_native.py): eval-mode BatchNorm (eps 1e-5; models/square.py:83-84 keeps the trunk's BN in eval
// mode) folded into each conv, weights transposed to the K-major blob layout of arch.py, conv_head / norm_head dropped, optional
// `...layer_scale.gamma` folded into pw_proj.  No GPU needed.
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>

#include "internal.h"
#include "arch_table.inc"

namespace {
struct View { const float* p; int64_t n; };
}

extern "C" int cv_square_pack_weights(const cv_named_tensor* tensors, int n_tensors, float* blob, size_t blob_floats) {
    CV_ARG(tensors != nullptr && blob != nullptr, "null argument");
    CV_ARG(blob_floats == (size_t)CV_BLOB_FLOATS, "blob buffer has the wrong number of floats (cv_weight_blob_floats())");
    std::unordered_map<std::string, View> sd;
    for (int i = 0; i < n_tensors; ++i) {
        CV_ARG(tensors[i].name != nullptr && (tensors[i].data != nullptr || tensors[i].numel == 0), "null tensor entry");
        sd[tensors[i].name] = View{tensors[i].data, tensors[i].numel};
    }
    auto need = [&](const std::string& key, int64_t numel, View* out) -> bool {
        auto it = sd.find(key);
        if (it == sd.end()) { cv_set_error("cv_square_pack_weights: missing tensor '%s'", key.c_str()); return false; }
        if (it->second.n != numel) {
            cv_set_error("cv_square_pack_weights: tensor '%s' has %lld elements, expected %lld", key.c_str(), (long long)it->second.n, (long long)numel);
            return false;
        }
        *out = it->second;
        return true;
    };
    memset(blob, 0, blob_floats * sizeof(float));
    for (int li = 0; li < CV_NUM_LAYERS; ++li) {
        const cv_layer_info& L = kLayers[li];
        const bool dw = L.kind == CV_KIND_DEPTHWISE;
        const int cin_g = dw ? 1 : L.cin, taps = L.k * L.k;
        const std::string pre = "backbone.", bn = pre + kLayerBnKey[li];
        View w, g, b, m, v;
        if (!need(pre + kLayerConvKey[li], (int64_t)L.cout * cin_g * taps, &w) || !need(bn + ".weight", L.cout, &g) || !need(bn + ".bias", L.cout, &b) ||
            !need(bn + ".running_mean", L.cout, &m) || !need(bn + ".running_var", L.cout, &v))
            return CV_ERR_ARG;
        // optional LayerScale on pw_proj (absent for the conv variants of MobileNetV4): key = <block>.layer_scale.gamma
        const float* gamma = nullptr;
        {
            const std::string key = kLayerConvKey[li];                      // e.g. blocks.2.1.pw_proj.conv.weight
            const size_t at = key.find(".pw_proj.");
            if (at != std::string::npos) {
                auto it = sd.find(pre + key.substr(0, at) + ".layer_scale.gamma");
                if (it != sd.end() && it->second.n == L.cout) gamma = it->second.p;
            }
        }
        float* wout = blob + L.w_offset;
        float* bout = blob + L.b_offset;
        for (int co = 0; co < L.cout; ++co) {
            double scale = (double)g.p[co] / std::sqrt((double)v.p[co] + 1e-5);
            double bias = (double)b.p[co] - (double)m.p[co] * scale;
            if (gamma) { scale = scale * (double)gamma[co]; bias = bias * (double)gamma[co]; }
            bout[co] = (float)bias;
            for (int ci = 0; ci < cin_g; ++ci)
                for (int t = 0; t < taps; ++t) {
                    const double val = (double)w.p[((int64_t)co * cin_g + ci) * taps + t] * scale;       // (O, I/g, ky, kx)
                    if (dw) wout[(int64_t)t * L.cout + co] = (float)val;                                 // [tap][c]
                    else wout[((int64_t)t * L.cin + ci) * L.cout + co] = (float)val;                    // [(ky,kx,ci)][co]
                }
        }
    }
    struct Part { const char* key; int64_t n; int64_t off; };
    const Part parts[] = {
        {"type_head.1.weight", 7 * 480, CV_OFF_HEAD_W},        {"color_head.1.weight", 3 * 480, CV_OFF_HEAD_W + 7 * 480},
        {"type_head.1.bias", 7, CV_OFF_HEAD_B},                {"color_head.1.bias", 3, CV_OFF_HEAD_B + 7},
        {"global_head.1.weight", 64LL * 30720, CV_OFF_GLOB_W}, {"global_head.1.bias", 64, CV_OFF_GLOB_B},
        {"turn_head.weight", 64, CV_OFF_TC_W},                 {"castling_head.weight", 4 * 64, CV_OFF_TC_W + 64},
        {"turn_head.bias", 1, CV_OFF_TC_B},                    {"castling_head.bias", 4, CV_OFF_TC_B + 1},
    };
    for (const Part& pt : parts) {
        View t;
        if (!need(pt.key, pt.n, &t)) return CV_ERR_ARG;
        memcpy(blob + pt.off, t.p, (size_t)pt.n * sizeof(float));
    }
    return CV_OK;
}
