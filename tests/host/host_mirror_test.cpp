// host_mirror_test.cpp -- GPU tests of the C++ host mirror (caffe_escoin_b200/host/escort_conv_layer.hpp), written in
// the shape of the reference's own convolution tests (src/caffe/test/test_convolution_layer.cpp): a layer is set up
// from a ConvolutionParameter, filled, run with Forward_gpu / Backward_gpu in SCONV mode and compared with the naive
// ground truth (there caffe_conv, :20-150; here the oracle's restatement of it, linked from oracle/libescort_oracle.so),
// EXPECT_NEAR 1e-4 like :231-265.  Cases mirror TestSimpleConvolution, TestSimpleConvolutionGroup, Test1x1Convolution,
// TestSobelConvolution (known answer) and the gradient checks (:709-855, here against the masked analytic gradient).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../caffe_escoin_b200/host/escort_conv_layer.hpp"
#include "../../caffe_escoin_b200/host/escort_dense_layers.hpp"

extern "C" {
void oracle_dense_conv(const float *bottom, int num, int Cin, int H, int W, const float *weights, int Cout, int group,
                       int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h,
                       int dilation_w, const float *bias, int fuse_relu, float *top);
void oracle_conv_backward(const float *bottom, const float *top_diff, int num, int Cin, int H, int W, const float *weights,
                          int Cout, int group, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h, int stride_w,
                          int dilation_h, int dilation_w, int mask_only, float *weight_diff, float *bias_diff,
                          float *bottom_diff);
}

using namespace escort_caffe;
static int g_fail = 0;
#define EXPECT(cond, ...)                 \
  do {                                    \
    if (!(cond)) {                        \
      ++g_fail;                           \
      printf("  FAILED %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                \
      printf("\n");                       \
    }                                     \
  } while (0)

static std::vector<float> d2h(const float *d, size_t n) {
  std::vector<float> h(n);
  cuda_check(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost), "d2h");
  return h;
}
static void h2d(float *d, const std::vector<float> &h) {
  cuda_check(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice), "h2d");
}
static double rel_l2(const std::vector<float> &a, const std::vector<float> &b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    num += ((double)a[i] - b[i]) * ((double)a[i] - b[i]);
    den += (double)b[i] * b[i];
  }
  return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}
static double max_abs(const std::vector<float> &a, const std::vector<float> &b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::fabs((double)a[i] - b[i]));
  return m;
}

struct Case {
  const char *name;
  int num, channels, height, width;
  ConvolutionParameter p;
  double sparsity;
  bool relu;
};

static void run_case(const Case &c) {
  printf("[ RUN      ] %s (%s)\n", c.name, c.relu ? "ConvolutionReLU" : "Convolution");
  const int fails_before = g_fail;
  std::mt19937 rng(1701);  // the reference's gradient-check seed, test_gradient_check_util.hpp:25
  std::normal_distribution<float> gauss(0.f, 1.f);
  std::uniform_real_distribution<float> uni(0.f, 1.f);
  ConvolutionLayer layer(c.p, c.relu);
  Blob bottom, top;
  bottom.Reshape({c.num, c.channels, c.height, c.width});
  layer.LayerSetUp(bottom.shape);
  top.Reshape(layer.top_shape());
  // fillers: gaussian weights pruned to the case's sparsity, constant-free bias, gaussian bottom
  std::vector<float> w(layer.blobs()[0].count()), b(c.p.bias_term ? c.p.num_output : 0), x(bottom.count());
  for (float &v : w) v = uni(rng) < c.sparsity ? 0.f : gauss(rng);
  for (float &v : b) v = 0.1f * gauss(rng);
  for (float &v : x) v = gauss(rng);
  h2d(layer.blobs()[0].mutable_gpu_data(), w);
  if (c.p.bias_term) h2d(layer.blobs()[1].mutable_gpu_data(), b);
  h2d(bottom.mutable_gpu_data(), x);

  Caffe::set_conv_mode(Caffe::SCONV);
  layer.WeightAlign();
  long nnz = 0, want = 0;
  for (int n : layer.nz_num()) nnz += n;
  for (float v : w) want += v != 0.f;
  EXPECT(nnz == want, "nz_num sums to %ld, dense weights hold %ld nonzeros", nnz, want);

  layer.Forward_gpu(bottom, top);
  cuda_check(cudaDeviceSynchronize(), "forward");
  std::vector<float> ref(top.count());
  const ConvolutionParameter &p = c.p;
  oracle_dense_conv(x.data(), c.num, c.channels, c.height, c.width, w.data(), p.num_output, p.group, p.kernel_h, p.kernel_w,
                    p.pad_h, p.pad_w, p.stride_h, p.stride_w, p.dilation_h, p.dilation_w, p.bias_term ? b.data() : nullptr,
                    c.relu ? 1 : 0, ref.data());
  std::vector<float> got = d2h(top.gpu_data(), top.count());
  EXPECT(max_abs(got, ref) < 1e-4 * 50 && rel_l2(got, ref) < 1e-4, "forward: max_abs %.3g rel_l2 %.3g (kernel %s)",
         max_abs(got, ref), rel_l2(got, ref), layer.kernel_name());

  if (!c.relu) {
    // gradient: top diff random; parameter diffs accumulate (start from a non-zero diff), bottom diff is overwritten
    std::vector<float> dy(top.count()), wd0(w.size()), bd0(b.size()), xd_garbage(x.size(), 123.f);
    for (float &v : dy) v = gauss(rng);
    for (float &v : wd0) v = 0.5f * gauss(rng);
    for (float &v : bd0) v = 0.5f * gauss(rng);
    h2d(top.mutable_gpu_diff(), dy);
    h2d(layer.blobs()[0].mutable_gpu_diff(), wd0);
    if (c.p.bias_term) h2d(layer.blobs()[1].mutable_gpu_diff(), bd0);
    h2d(bottom.mutable_gpu_diff(), xd_garbage);
    layer.Backward_gpu(top, true, bottom);
    cuda_check(cudaDeviceSynchronize(), "backward");
    std::vector<float> wd_ref = wd0, bd_ref = bd0, xd_ref(x.size(), 0.f);
    oracle_conv_backward(x.data(), dy.data(), c.num, c.channels, c.height, c.width, w.data(), p.num_output, p.group,
                         p.kernel_h, p.kernel_w, p.pad_h, p.pad_w, p.stride_h, p.stride_w, p.dilation_h, p.dilation_w, 1,
                         wd_ref.data(), p.bias_term ? bd_ref.data() : nullptr, xd_ref.data());
    std::vector<float> wd = d2h(layer.blobs()[0].gpu_diff(), w.size()), xd = d2h(bottom.gpu_diff(), x.size());
    EXPECT(rel_l2(wd, wd_ref) < 1e-4, "weight diff (masked, accumulated): rel_l2 %.3g", rel_l2(wd, wd_ref));
    EXPECT(rel_l2(xd, xd_ref) < 1e-4, "bottom diff (overwritten): rel_l2 %.3g", rel_l2(xd, xd_ref));
    if (c.p.bias_term) {
      std::vector<float> bd = d2h(layer.blobs()[1].gpu_diff(), b.size());
      EXPECT(rel_l2(bd, bd_ref) < 1e-4, "bias diff (accumulated): rel_l2 %.3g", rel_l2(bd, bd_ref));
    }
    // param_propagate_down_[0] = false leaves the weight diff untouched (conv_layer.cu:56)
    layer.param_propagate_down()[0] = false;
    h2d(layer.blobs()[0].mutable_gpu_diff(), wd0);
    layer.Backward_gpu(top, false, bottom);
    cuda_check(cudaDeviceSynchronize(), "backward 2");
    EXPECT(max_abs(d2h(layer.blobs()[0].gpu_diff(), w.size()), wd0) == 0.0, "weight diff touched although param_propagate_down_[0] is false");
  }
  printf("[ %s ] %s\n", g_fail == fails_before ? "      OK" : " FAILED ", c.name);
}

static void sobel_known_answer() {
  // TestSobelConvolution (:498-589): 3x3 Sobel filter over a constant-gradient image has a closed-form answer
  printf("[ RUN      ] SobelKnownAnswer\n");
  const int fails_before = g_fail;
  ConvolutionParameter p;
  p.num_output = 1; p.bias_term = false; p.kernel_h = p.kernel_w = 3; p.stride_h = p.stride_w = 1;
  ConvolutionLayer layer(p);
  Blob bottom, top;
  bottom.Reshape({2, 1, 8, 9});
  layer.LayerSetUp(bottom.shape);
  top.Reshape(layer.top_shape());
  std::vector<float> x(bottom.count());
  for (int n = 0; n < 2; ++n)
    for (int y = 0; y < 8; ++y)
      for (int xx = 0; xx < 9; ++xx) x[(n * 8 + y) * 9 + xx] = 3.f * xx + 0.5f * y + n;  // d/dx = 3
  h2d(bottom.mutable_gpu_data(), x);
  h2d(layer.blobs()[0].mutable_gpu_data(), {-1, 0, 1, -2, 0, 2, -1, 0, 1});  // the zeros are pruned by the pack
  Caffe::set_conv_mode(Caffe::SCONV_PAR);
  layer.WeightAlign();
  EXPECT(layer.nz_num()[0] == 6, "Sobel filter should pack to 6 nonzeros, got %d", layer.nz_num()[0]);
  layer.Forward_gpu(bottom, top);
  cuda_check(cudaDeviceSynchronize(), "forward");
  for (float v : d2h(top.gpu_data(), top.count())) EXPECT(std::fabs(v - 24.f) < 1e-4, "Sobel response %.6f != 24", v);
  printf("[ %s ] SobelKnownAnswer\n", g_fail == fails_before ? "      OK" : " FAILED ");
}

static void error_behaviour() {
  printf("[ RUN      ] ErrorBehaviour\n");
  const int fails_before = g_fail;
  ConvolutionParameter p;
  p.num_output = 4; p.kernel_h = p.kernel_w = 3;
  ConvolutionLayer layer(p);
  Blob bottom, top;
  bottom.Reshape({1, 2, 6, 6});
  layer.LayerSetUp(bottom.shape);
  top.Reshape(layer.top_shape());
  bool threw = false;
  Caffe::set_conv_mode(Caffe::SCONV);
  try { layer.Forward_gpu(bottom, top); } catch (const std::runtime_error &) { threw = true; }
  EXPECT(threw, "Forward_gpu before WeightAlign must fail loudly");
  layer.WeightAlign();  // all-zero weights: nnz == 0 is legal
  layer.Forward_gpu(bottom, top);
  threw = false;
  Caffe::set_conv_mode(Caffe::LOWERED_GEMM);
  try { layer.Forward_gpu(bottom, top); } catch (const std::runtime_error &) { threw = true; }
  EXPECT(threw, "dense conv_mode must not silently run the sparse path");
  printf("[ %s ] ErrorBehaviour\n", g_fail == fails_before ? "      OK" : " FAILED ");
}

// ---- the layers the reference keeps dense (f1), TF32 on tcgen05: EXPECT_NEAR at the TF32 bar (2e-3 relative L2) ----
static void inner_product_forward(int num, int c, int h, int w, int num_output, bool bias) {  // test_inner_product_layer.cpp TestForward
  printf("[ RUN      ] InnerProductForward_%dx%dx%dx%d_to_%d\n", num, c, h, w, num_output);
  const int fails_before = g_fail;
  std::mt19937 rng(1701);
  std::uniform_real_distribution<float> uni(-1.f, 1.f);
  InnerProductParameter ip;
  ip.num_output = num_output;
  ip.bias_term = bias;
  InnerProductLayer layer(ip);
  Blob bottom, top;
  bottom.Reshape({num, c, h, w});
  layer.LayerSetUp(bottom.shape);
  top.Reshape(layer.top_shape());
  const int K = c * h * w;
  std::vector<float> wv(layer.blobs()[0].count()), b(bias ? num_output : 0), x(bottom.count());
  for (float &v : wv) v = uni(rng) / std::sqrt((float)K);
  for (float &v : b) v = uni(rng);
  for (float &v : x) v = uni(rng);
  h2d(layer.blobs()[0].mutable_gpu_data(), wv);
  if (bias) h2d(layer.blobs()[1].mutable_gpu_data(), b);
  h2d(bottom.mutable_gpu_data(), x);
  layer.Forward_gpu(bottom, top);
  cuda_check(cudaDeviceSynchronize(), "inner product forward");
  std::vector<float> ref(top.count());
  // an inner product is a 1x1 convolution of a K-channel 1x1 image
  oracle_dense_conv(x.data(), num, K, 1, 1, wv.data(), num_output, 1, 1, 1, 0, 0, 1, 1, 1, 1, bias ? b.data() : nullptr, 0, ref.data());
  std::vector<float> got = d2h(top.gpu_data(), top.count());
  EXPECT(rel_l2(got, ref) < 2e-3, "inner product: rel_l2 %.3g", rel_l2(got, ref));
  printf("[ %s ] InnerProductForward\n", g_fail == fails_before ? "      OK" : " FAILED ");
}

static void esc_convolution_forward(const char *name, int num, int channels, int hw, const ConvolutionParameter &p, bool residual) {
  printf("[ RUN      ] EscConvolution_%s\n", name);
  const int fails_before = g_fail;
  std::mt19937 rng(1701);
  std::uniform_real_distribution<float> uni(-1.f, 1.f);
  EscConvolutionLayer layer(p, /*fuse_relu=*/residual);
  Blob bottom, top, shortcut;
  bottom.Reshape({num, channels, hw, hw});
  layer.LayerSetUp(bottom.shape);
  top.Reshape(layer.top_shape());
  shortcut.Reshape(layer.top_shape());
  std::vector<float> wv(layer.blobs()[0].count()), b(p.bias_term ? p.num_output : 0), x(bottom.count()), r(top.count());
  const float scale = 1.f / std::sqrt((float)(channels * p.kernel_h * p.kernel_w));
  for (float &v : wv) v = uni(rng) * scale;
  for (float &v : b) v = uni(rng);
  for (float &v : x) v = uni(rng);
  for (float &v : r) v = uni(rng);
  h2d(layer.blobs()[0].mutable_gpu_data(), wv);
  if (p.bias_term) h2d(layer.blobs()[1].mutable_gpu_data(), b);
  h2d(bottom.mutable_gpu_data(), x);
  h2d(shortcut.mutable_gpu_data(), r);
  layer.Forward_gpu(bottom, top, residual ? &shortcut : nullptr);
  cuda_check(cudaDeviceSynchronize(), "dense conv forward");
  std::vector<float> ref(top.count());
  oracle_dense_conv(x.data(), num, channels, hw, hw, wv.data(), p.num_output, 1, p.kernel_h, p.kernel_w, p.pad_h, p.pad_w, p.stride_h,
                    p.stride_w, p.dilation_h, p.dilation_w, p.bias_term ? b.data() : nullptr, 0, ref.data());
  if (residual)
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = std::max(ref[i] + r[i], 0.f);  // Eltwise SUM + ReLU
  std::vector<float> got = d2h(top.gpu_data(), top.count());
  EXPECT(rel_l2(got, ref) < 2e-3, "dense convolution: rel_l2 %.3g", rel_l2(got, ref));
  printf("[ %s ] EscConvolution_%s\n", g_fail == fails_before ? "      OK" : " FAILED ", name);
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    printf("no CUDA device: the host mirror has no CPU fallback\n");
    return 77;
  }
  auto P = [](int no, int k, int s, int pad, int g, bool bias, int dil = 1) {
    ConvolutionParameter p;
    p.num_output = no; p.kernel_h = p.kernel_w = k; p.stride_h = p.stride_w = s; p.pad_h = p.pad_w = pad; p.group = g;
    p.bias_term = bias; p.dilation_h = p.dilation_w = dil;
    return p;
  };
  const Case cases[] = {
      {"SimpleConvolution_2x3x6x4_k3_s2", 2, 3, 6, 4, P(4, 3, 2, 0, 1, true), 0.5, false},     // :231-265
      {"SimpleConvolutionGroup", 2, 6, 6, 4, P(6, 3, 2, 0, 3, true), 0.5, false},              // :470-496
      {"Convolution1x1", 2, 3, 6, 4, P(4, 1, 1, 0, 1, true), 0.3, false},                      // :384-414
      {"DilatedConvolution", 2, 3, 9, 9, P(4, 3, 1, 0, 1, true, 2), 0.5, false},               // :267-309
      {"AlexNetConv3Like_13x13_pad1", 5, 32, 13, 13, P(48, 3, 1, 1, 1, true), 0.88, false},
      {"AlexNetConv2Like_27x27_k5_g2", 3, 16, 27, 27, P(32, 5, 1, 2, 2, true), 0.85, false},
      {"ResNetLike_28x28_nobias_tma", 4, 16, 28, 28, P(16, 3, 1, 1, 1, false), 0.7, false},
      {"ResNetLike_56x56_nobias_tma", 3, 8, 56, 56, P(8, 3, 1, 1, 1, false), 0.7, false},
      {"ConvolutionReLU_14x14", 4, 24, 14, 14, P(40, 3, 1, 1, 1, true), 0.75, true},            // conv_relu_layer.cu
      {"Stride2_28x28", 2, 8, 28, 28, P(16, 3, 2, 1, 1, true), 0.6, false},
  };
  for (const Case &c : cases) run_case(c);
  sobel_known_answer();
  error_behaviour();
  inner_product_forward(2, 3, 4, 5, 10, true);      // test_inner_product_layer.cpp: bottom 2x3x4x5, num_output 10 (K = 60)
  inner_product_forward(130, 9, 2, 2, 257, false);  // ragged in every dimension of the 128-row tiles
  esc_convolution_forward("Conv1Like_k7_s2_p3", 2, 3, 30, P(16, 7, 2, 3, 1, true), false);
  esc_convolution_forward("Pointwise_implicit_gemm", 3, 36, 14, P(72, 1, 1, 0, 1, true), false);
  esc_convolution_forward("Branch2c_residual_relu", 2, 32, 12, P(96, 1, 1, 0, 1, true), true);
  printf(g_fail ? "[  FAILED  ] %d expectation(s)\n" : "[  PASSED  ] all host-mirror tests\n", g_fail);
  return g_fail ? 1 : 0;
}
