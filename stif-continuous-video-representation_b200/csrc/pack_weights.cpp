// pack_weights.cpp -- fold the 26 reference tensors into the hoisted formulation (host, float64).
//
// Reference layer shapes: Sakuya_arch_test.py:306-311; concat orders :399 (201 = latent 192, frames 6,
// rel_y, rel_x, t), :418 (263 = HRfeat 64, latent 192, frames 6, t), :455 (525 = HRfeat@g1 64,
// HRfeat@g2 64, latent@g1 192, latent@g2 192, frames@g1 6, frames@g2 6, t).
#include "stif_internal.h"

namespace stif {
namespace {

constexpr double kOmega = 30.0;  // SIREN.py:24,45 (first_omega_0 = hidden_omega_0 = 30)

struct Mat {
  const float* p;
  int rows, cols;
  double at(int r, int c) const { return (double)p[(size_t)r * cols + c]; }
};

void scaled_copy(std::vector<float>& dst, const float* src, size_t n, double s) {
  dst.resize(n);
  for (size_t i = 0; i < n; ++i) dst[i] = (float)(s * (double)src[i]);
}

// dst[r, :] = s * src[r, c0:c1]  (appended at column offset dcol of a [rows, dcols] matrix)
void put_cols(std::vector<float>& dst, int dcols, int drow0, int dcol, const Mat& m, int c0, int c1, double s) {
  for (int r = 0; r < m.rows; ++r)
    for (int c = c0; c < c1; ++c) dst[(size_t)(drow0 + r) * dcols + dcol + (c - c0)] = (float)(s * m.at(r, c));
}

void column(std::vector<float>& dst, const Mat& m, int c, double s) {
  dst.resize(m.rows);
  for (int r = 0; r < m.rows; ++r) dst[r] = (float)(s * m.at(r, c));
}

}  // namespace

void fold_weights(const float* const* t, FoldedWeights& o) {
  // ABI order: feat 0..7, flow 8..15, encode 16..25 (weight, bias pairs)
  Mat Wf0{t[0], 64, 201}, Wf1{t[2], 64, 64}, Wf2{t[4], 256, 64}, Wf3{t[6], 64, 256};
  const float *bf0 = t[1], *bf1 = t[3], *bf2 = t[5], *bf3 = t[7];
  Mat Wl0{t[8], 64, 263}, Wl1{t[10], 64, 64}, Wl2{t[12], 256, 64}, Wl3{t[14], 4, 256};
  const float *bl0 = t[9], *bl1 = t[11], *bl2 = t[13], *bl3 = t[15];
  Mat We0{t[16], 64, 525}, We1{t[18], 64, 64}, We2{t[20], 256, 64}, We3{t[22], 256, 256}, We4{t[24], 3, 256};
  const float *be0 = t[17], *be1 = t[19], *be2 = t[21], *be3 = t[23], *be4 = t[25];

  // latent projection [256,198]
  o.w_tab.assign((size_t)256 * 198, 0.f);
  put_cols(o.w_tab, 198, 0, 0, Wf0, 0, 198, kOmega);        // TA : latent+frames columns of feat_imnet L0
  put_cols(o.w_tab, 198, 64, 0, Wl0, 64, 262, kOmega);      // TB : latent+frames columns of flow_imnet L0
  put_cols(o.w_tab, 198, 128, 0, We0, 128, 320, kOmega);    // TE1: latent@g1 ...
  put_cols(o.w_tab, 198, 128, 192, We0, 512, 518, kOmega);  //      frames@g1
  put_cols(o.w_tab, 198, 192, 0, We0, 320, 512, kOmega);    // TE2: latent@g2 ...
  put_cols(o.w_tab, 198, 192, 192, We0, 518, 524, kOmega);  //      frames@g2
  // decoding_test variant: the bilinear frame gathers read the x4-upsampled pair, so the frame columns of TB | TE1 | TE2
  // leave the LR table and become a [192,6] map applied on the 4H x 4W grid (TA keeps them: its gather is the LR nearest)
  o.w_tab_lat = o.w_tab;
  o.w_up.assign((size_t)192 * 6, 0.f);
  for (int r = 64; r < 256; ++r)
    for (int c = 0; c < 6; ++c) {
      o.w_up[(size_t)(r - 64) * 6 + c] = o.w_tab[(size_t)r * 198 + 192 + c];
      o.w_tab_lat[(size_t)r * 198 + 192 + c] = 0.f;
    }

  o.a_rel.resize(128);
  for (int r = 0; r < 64; ++r) {
    o.a_rel[r * 2 + 0] = (float)(kOmega * Wf0.at(r, 198));
    o.a_rel[r * 2 + 1] = (float)(kOmega * Wf0.at(r, 199));
  }
  column(o.a_t, Wf0, 200, kOmega);
  scaled_copy(o.a_b, bf0, 64, kOmega);
  scaled_copy(o.f1_w, Wf1.p, 64 * 64, kOmega);
  scaled_copy(o.f1_b, bf1, 64, kOmega);
  scaled_copy(o.f2_w, Wf2.p, 256 * 64, kOmega);
  scaled_copy(o.f2_b, bf2, 256, kOmega);

  // composed last layer of feat_imnet: rows F (flow L0 cols 0:64), Q1 (encode L0 cols 0:64), Q2 (cols 64:128)
  o.f3_w.assign((size_t)192 * 256, 0.f);
  o.f3_b.assign(192, 0.f);
  for (int part = 0; part < 3; ++part) {
    const Mat& src = part == 0 ? Wl0 : We0;
    int c0 = part == 2 ? 64 : 0;
    for (int r = 0; r < 64; ++r) {
      for (int k = 0; k < 256; ++k) {
        double acc = 0.0;
        for (int j = 0; j < 64; ++j) acc += src.at(r, c0 + j) * Wf3.at(j, k);
        o.f3_w[(size_t)(part * 64 + r) * 256 + k] = (float)(kOmega * acc);
      }
      double accb = 0.0;
      for (int j = 0; j < 64; ++j) accb += src.at(r, c0 + j) * (double)bf3[j];
      o.f3_b[part * 64 + r] = (float)(kOmega * accb);
    }
  }

  column(o.b_t, Wl0, 262, kOmega);
  scaled_copy(o.b_b, bl0, 64, kOmega);
  scaled_copy(o.l1_w, Wl1.p, 64 * 64, kOmega);
  scaled_copy(o.l1_b, bl1, 64, kOmega);
  scaled_copy(o.l2_w, Wl2.p, 256 * 64, kOmega);
  scaled_copy(o.l2_b, bl2, 256, kOmega);
  scaled_copy(o.l3_w, Wl3.p, 4 * 256, 1.0);
  scaled_copy(o.l3_b, bl3, 4, 1.0);

  column(o.e_t, We0, 524, kOmega);
  scaled_copy(o.e_b, be0, 64, kOmega);
  scaled_copy(o.e1_w, We1.p, 64 * 64, kOmega);
  scaled_copy(o.e1_b, be1, 64, kOmega);
  scaled_copy(o.e2_w, We2.p, 256 * 64, kOmega);
  scaled_copy(o.e2_b, be2, 256, kOmega);
  scaled_copy(o.e3_w, We3.p, 256 * 256, kOmega);
  scaled_copy(o.e3_b, be3, 256, kOmega);
  scaled_copy(o.e4_w, We4.p, 3 * 256, 1.0);
  scaled_copy(o.e4_b, be4, 3, 1.0);
}

}  // namespace stif
