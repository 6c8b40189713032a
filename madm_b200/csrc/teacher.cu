// Teacher post-processing of MADM's self-training step (reference modeling/meta_arch/mtmadise.py:337-352,
// utils/dacs_transforms.py:87-112): everything between the EMA head's logits and the mixed pseudo-labels, fused and without the
// reference's host round trips (`pseudo_label.cpu()`, `torch.sum(...).item()`).
//   pseudo_label_kernel   bilinear upsample of the logits to the image size (align_corners=False) -> softmax over classes ->
//                         max probability + argmax per pixel, and the count of pixels whose confidence reaches the threshold
//   pseudo_weight_kernel  pseudo_weight = (count / pixels) everywhere, 0 in the top `ignore_top` rows (pl_crop)
//   class_mask_kernel     generate_class_mask: 1 where the label is one of the chosen classes
//   one_mix_*_kernel      DACS mixing  mask * a + (1 - mask) * b  for labels (int64) and pixel weights (fp32)
// All HBM-bound (the 128x128 logits are L2-resident); the count uses integer atomics only, so results are run-to-run identical.
#include "kernels.h"

namespace madm {

static constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(256) pseudo_label_kernel(const float* __restrict__ logits, int C, int h, int w, int H, int W, float sy,
                                                           float sx, float threshold, long total, int64_t* __restrict__ label,
                                                           float* __restrict__ prob, int* __restrict__ count) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  int confident = 0;
  if (i < total) {
    const int x = int(i % W), y = int((i / W) % H), b = int(i / (long(W) * H));
    // F.interpolate(mode='bilinear', align_corners=False): src = (dst + 0.5) * scale - 0.5, clamped at 0
    const float fy = fmaxf((y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min(int(fy), h - 1), x0 = min(int(fx), w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - float(y0), lx = fx - float(x0), hy = 1.f - ly, hx = 1.f - lx;
    const float* base = logits + size_t(b) * C * h * w;
    const int o00 = y0 * w + x0, o01 = y0 * w + x1, o10 = y1 * w + x0, o11 = y1 * w + x1;
    float v[kMaxClasses];
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) {
      if (c < C) {
        const float* p = base + size_t(c) * h * w;
        // same association as ATen's upsample_bilinear2d: hy * (hx * a + lx * b) + ly * (hx * c + lx * d)
        v[c] = hy * (hx * __ldg(p + o00) + lx * __ldg(p + o01)) + ly * (hx * __ldg(p + o10) + lx * __ldg(p + o11));
        if (v[c] > best) { best = v[c]; arg = c; }  // first maximum, like torch.max
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) sum += expf(v[c] - best);
    const float pmax = 1.0f / sum;  // softmax probability of the arg-max class
    label[i] = arg;
    prob[i] = pmax;
    confident = pmax >= threshold ? 1 : 0;
  }
  const int n = __syncthreads_count(confident);
  if (threadIdx.x == 0 && n) atomicAdd(count, n);
}

__global__ void pseudo_weight_kernel(const int* __restrict__ count, long total, int H, int W, int ignore_top, float* __restrict__ weight) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int y = int((i / W) % H);
  const float ratio = float(double(*count) / double(total));  // python float division, then * torch.ones(float32)
  weight[i] = y < ignore_top ? 0.f : ratio;
}

const char* pseudo_labels(const float* logits, int B, int C, int h, int w, int H, int W, float threshold, int ignore_top, int64_t* label,
                          float* prob, float* weight, int* count, cudaStream_t st) {
  if (C < 1 || C > kMaxClasses) return "pseudo_labels: 1..32 classes supported";
  if (B < 1 || h < 1 || w < 1 || H < 1 || W < 1) return "pseudo_labels: empty input";
  const long total = long(B) * H * W;
  if (cudaMemsetAsync(count, 0, sizeof(int), st) != cudaSuccess) return "pseudo_labels: cudaMemsetAsync failed";
  pseudo_label_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(logits, C, h, w, H, W, float(h) / float(H), float(w) / float(W), threshold,
                                                                    total, label, prob, count);
  if (cudaGetLastError() != cudaSuccess) return "pseudo_label launch failed";
  if (weight) {
    pseudo_weight_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(count, total, H, W, ignore_top, weight);
    if (cudaGetLastError() != cudaSuccess) return "pseudo_weight launch failed";
  }
  return nullptr;
}

// generate_class_mask (dacs_transforms.py:98-103): mask[p] = sum_k (label[p] == classes[k])  (0/1 for distinct classes)
__global__ void class_mask_kernel(const int64_t* __restrict__ label, long n, const int64_t* __restrict__ classes, int k, int64_t* __restrict__ mask) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t l = label[i];
  int64_t m = 0;
  for (int j = 0; j < k; ++j) m += (l == __ldg(classes + j)) ? 1 : 0;
  mask[i] = m;
}
const char* class_mask(const int64_t* label, long n, const int64_t* classes, int k, int64_t* mask, cudaStream_t st) {
  if (n < 1 || k < 0) return "class_mask: bad size";
  class_mask_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(label, n, classes, k, mask);
  return cudaGetLastError() == cudaSuccess ? nullptr : "class_mask launch failed";
}

// one_mix (dacs_transforms.py:106-112): out = mask * a + (1 - mask) * b, for int64 labels and fp32 pixel weights in one pass
__global__ void one_mix_kernel(const int64_t* __restrict__ mask, long n, const int64_t* __restrict__ la, const int64_t* __restrict__ lb,
                               int64_t* __restrict__ lout, const float* __restrict__ wa, const float* __restrict__ wb, float* __restrict__ wout) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t m = mask[i];
  if (lout) lout[i] = m * la[i] + (1 - m) * lb[i];
  if (wout) wout[i] = float(m) * wa[i] + float(1 - m) * wb[i];
}
const char* one_mix(const int64_t* mask, long n, const int64_t* la, const int64_t* lb, int64_t* lout, const float* wa, const float* wb,
                    float* wout, cudaStream_t st) {
  if (n < 1) return "one_mix: bad size";
  if ((lout && (!la || !lb)) || (wout && (!wa || !wb))) return "one_mix: missing operand";
  one_mix_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(mask, n, la, lb, lout, wa, wb, wout);
  return cudaGetLastError() == cudaSuccess ? nullptr : "one_mix launch failed";
}

// ------------------------------------------------------------------ sliding-window merge (feature_extractor.py:254-275)
// feats [nwin*n, C, hf, wf] (window-major: crop wi of image b at row wi*n + b) -> out [n, C, Hf, Wf] = sum over the windows that
// cover a pixel / their count.  Gather form: each output pixel adds its windows in window order (the reference's `+=` order) and
// divides by the count, so there are no atomics and the result is bit-identical to the sequential accumulate.
__global__ void slide_merge_kernel(const float* __restrict__ feats, int nwin, int n, int C, int hf, int wf, const int* __restrict__ wins,
                                   int Hf, int Wf, long total, float* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = int(i % Wf), y = int((i / Wf) % Hf);
  const long bc = i / (long(Wf) * Hf);
  const int c = int(bc % C), b = int(bc / C);
  float acc = 0.f;
  int cnt = 0;
  for (int wi = 0; wi < nwin; ++wi) {
    const int yy = y - wins[2 * wi], xx = x - wins[2 * wi + 1];
    if (yy >= 0 && yy < hf && xx >= 0 && xx < wf) {
      acc += feats[((size_t(wi) * n + b) * C + c) * hf * wf + size_t(yy) * wf + xx];
      ++cnt;
    }
  }
  out[i] = acc / float(cnt);  // cnt == 0 cannot happen for the reference's window grids (inf/nan like the reference's 0/0 otherwise)
}
const char* slide_merge(const float* feats, int nwin, int n, int C, int hf, int wf, const int* wins, int Hf, int Wf, float* out, cudaStream_t st) {
  if (nwin < 1 || n < 1 || C < 1 || hf < 1 || wf < 1 || Hf < hf || Wf < wf) return "slide_merge: bad geometry";
  const long total = long(n) * C * Hf * Wf;
  slide_merge_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(feats, nwin, n, C, hf, wf, wins, Hf, Wf, total, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "slide_merge launch failed";
}

}  // namespace madm
