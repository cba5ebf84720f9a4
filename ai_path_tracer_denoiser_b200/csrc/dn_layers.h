// The 28 convolutions of training/recurrent_autoencoder_model.py in execution order (:129-140), with the
// state_dict keys of each conv / BatchNorm pair - mirrors weights.py:conv_layers().
#pragma once
#include <string>
#include <vector>

enum DnKind { DN_L1 = 0, DN_L2A = 1, DN_L2B = 2, DN_DEC1 = 3, DN_DEC2 = 4 };
struct DnLayerSpec {
    std::string name, conv_key, bn_key;
    DnKind kind;
    int level;        // output resolution = padded frame >> level
    int cin0, cin1;   // real channel counts of the (up to two) concatenated sources
    int cout;
    bool lrelu_first; // encoder layer2's first conv applies LeakyReLU before BatchNorm (model.py:30-32)
};

inline std::vector<DnLayerSpec> dn_layer_specs() {
    static const int ENC[5][2] = {{10, 32}, {32, 43}, {43, 57}, {57, 76}, {76, 101}};   // model.py:98-107
    static const int DEC[5][2] = {{101, 76}, {76, 57}, {57, 43}, {43, 32}, {32, 3}};    // :111-115 (decoder5 .. decoder1)
    std::vector<DnLayerSpec> L;
    for (int k = 1; k <= 5; ++k) {
        const std::string p = "encoder" + std::to_string(k) + ".0.", n = "enc" + std::to_string(k);
        const int ci = ENC[k - 1][0], co = ENC[k - 1][1];
        L.push_back({n + ".l1", p + "layer1.0", p + "layer1.1", DN_L1, k - 1, ci, 0, co, false});       // :23-27
        L.push_back({n + ".l2a", p + "layer2.0", p + "layer2.2", DN_L2A, k - 1, co, co, co, true});     // :30-32
        L.push_back({n + ".l2b", p + "layer2.3", p + "layer2.4", DN_L2B, k - 1, co, 0, co, false});     // :33-35
    }
    L.push_back({"bott.l1", "bottleneck.layer1.0", "bottleneck.layer1.1", DN_L1, 5, 101, 0, 101, false});   // :50-54
    L.push_back({"bott.l2a", "bottleneck.layer2.0", "bottleneck.layer2.1", DN_L2A, 5, 101, 101, 101, false});
    L.push_back({"bott.l2b", "bottleneck.layer2.3", "bottleneck.layer2.4", DN_L2B, 5, 101, 0, 101, false});
    for (int i = 0; i < 5; ++i) {
        const int k = 5 - i, ci = DEC[i][0], co = DEC[i][1];
        const std::string p = "decoder" + std::to_string(k) + ".layer1.", n = "dec" + std::to_string(k);
        L.push_back({n + ".c1", p + "1", p + "2", DN_DEC1, k - 1, ci, ci, co, false});                   // :40-43
        L.push_back({n + ".c2", p + "4", p + "5", DN_DEC2, k - 1, co, 0, co, false});                    // :44-46
    }
    return L;
}
