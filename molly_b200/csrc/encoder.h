// The encoder handle behind `molly_encoder_t` (private to the library: abi.cu builds it, train.cu reads it).
#pragma once
#include <vector>

#include "../../include/molly_b200.h"
#include "kernels.h"

struct molly_encoder {
    molly_encoder_config cfg;
    molly_encoder_weights w;                     // scalar members + pointers into the vectors below
    std::vector<const float*> ln1_w, ln1_b, b_qkv, b_o, ln2_w, ln2_b, b_ffn1, b_ffn2;
    std::vector<const void*> w_qkv, w_o, w_ffn1, w_ffn2;
    std::vector<CUtensorMap> tm_wqkv, tm_wo, tm_w1, tm_w2;   // weight (B operand) tensor maps, built once
    CUtensorMap tm_wproj;
    int ffn1_n;                                  // F (gelu) or 2F (glu)
    float q_scale;                               // head_dim^-1/2
    // activation tensor maps depend on (workspace, n_seq, K): cached for the last plan
    struct Plan {
        void* ws = nullptr;
        int n_seq = 0, k = 0;
        void* final_out = nullptr;
        CUtensorMap tm_xn, tm_attn, tm_mid, tm_final;   // A operands
        molly::AttnMaps tm_qkv;                                // attention input
        CUtensorMap tc_qkv, tc_x, tc_mid;                       // GEMM outputs (tc_x also feeds the residual loads)
    } plan;
};

