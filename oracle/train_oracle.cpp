// train_oracle.cpp — CPU oracle of the training step.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
// bench.py's CPU-baseline legs may load this; nothing under alphagpu_b200/ does.
//
// PARITY UNPINNED: the reference's training step is Flux 0.12.6 + Zygote 0.6.19 + cuBLAS through CUDA.jl 3.3.5 (Manifest.toml),
// none of which is under /root/reference and none of which can run here (no Julia).  This file restates the published
// algorithm from the reference's own call sites:
//   networkf forward, training branch      DenseNet.jl:161-186 (resnets: relu.(x .+ c1(x)), c1 = Dense(n,n,relu), :33-39)
//   lossTot                                train.jl:12-15      logitcrossentropy(p,y1) + mse(v,y2) + 0.001f0*mse(f,y3)
//   custom_train! / Flux.update!           train.jl:128-162
//   Optimiser(ADAM(lr), WeightDecay(1e-4)) train.jl:50         Flux 0.12.6 optimisers.jl: apply!(ADAM) then apply!(WeightDecay)
// with Zygote's pullbacks written out by hand (relu' = x>0, sigma' = y(1-y), tanh' = 1-y^2, dlogsoftmax = D - sum(D) softmax).
// The floating-point ORDER is a specification shared with alphagpu_b200/csrc/train.cu (DESIGN.md "training step"): fma chains
// ascending in k, weight gradients per 256-sample slice then slices in order, Adam per element in Float64.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" void orc_expf(const float* x, int64_t n, float* y);   // the canonical exp of oracle.cpp

namespace {

constexpr int KSLICE = 256;

inline float c_expf(float x) { float y; orc_expf(&x, 1, &y); return y; }
inline float c_sigmoidf(float x) { float t = c_expf(x < 0.f ? x : -x); float d = 1.0f + t; return x >= 0.f ? 1.0f / d : t / d; }
inline float c_tanhf(float x) {
  const float ax = x < 0.f ? -x : x;
  float t;
  if (ax < 0.25f) {
    const float s = ax * ax;
    float p = 2.18694885361552e-2f;
    p = p * s + -5.39682539682540e-2f;
    p = p * s + 1.33333333333333e-1f;
    p = p * s + -3.33333333333333e-1f;
    p = p * s + 1.0f;
    t = ax * p;
  } else if (ax > 9.0f) {
    t = 1.0f;
  } else {
    const float e = c_expf(2.0f * ax);
    t = 1.0f - 2.0f / (e + 1.0f);
  }
  return x < 0.f ? -t : t;
}
inline float relu(float x) { return x > 0.f ? x : 0.f; }

struct Trainer {
  int in, n, k, A, FS, NH;
  double lr, b1, b2, eps, wd;
  float fweight;
  int64_t P, off_res, off_heads, off_bias;
  std::vector<float> params, grads, m, v;
  double betap[2];
};

// C[i][j] = sum_{k in [k0,k1)} A(i,k) B(k,j), fma chain ascending from +0
template <class FA, class FB>
inline float dot(FA a, FB b, int k0, int k1) {
  float acc = 0.f;
  for (int k = k0; k < k1; k++) acc = fmaf(a(k), b(k), acc);
  return acc;
}

}  // namespace

extern "C" {

void* orc_trainer_create(int in, int n, int k, int A, int FS, double lr, double b1, double b2, double eps, double wd, float fweight) {
  Trainer* t = new Trainer();
  t->in = in; t->n = n; t->k = k; t->A = A; t->FS = FS; t->NH = A + 1 + FS;
  t->lr = lr; t->b1 = b1; t->b2 = b2; t->eps = eps; t->wd = wd; t->fweight = fweight;
  t->off_res = (int64_t)n * in;
  t->off_heads = t->off_res + (int64_t)k * n * n;
  t->off_bias = t->off_heads + (int64_t)t->NH * n;
  t->P = t->off_bias + t->NH;
  t->params.assign(t->P, 0.f); t->grads.assign(t->P, 0.f); t->m.assign(t->P, 0.f); t->v.assign(t->P, 0.f);
  t->betap[0] = b1; t->betap[1] = b2;
  return t;
}
void orc_trainer_destroy(void* h) { delete (Trainer*)h; }
int64_t orc_trainer_count(void* h) { return ((Trainer*)h)->P; }
// which: 0 params, 1 grads, 2 first moments, 3 second moments (flat: base, res, heads packed (NH x n) column-major, head biases)
int orc_trainer_get(void* h, int which, float* out) {
  Trainer* t = (Trainer*)h;
  const std::vector<float>* src[4] = {&t->params, &t->grads, &t->m, &t->v};
  if (which < 0 || which > 3) return -1;
  memcpy(out, src[which]->data(), sizeof(float) * t->P);
  return 0;
}
int orc_trainer_set(void* h, int which, const float* in) {
  Trainer* t = (Trainer*)h;
  std::vector<float>* dst[4] = {&t->params, &t->grads, &t->m, &t->v};
  if (which < 0 || which > 3) return -1;
  memcpy(dst[which]->data(), in, sizeof(float) * t->P);
  return 0;
}
void orc_trainer_reset_opt(void* h) {
  Trainer* t = (Trainer*)h;
  std::fill(t->m.begin(), t->m.end(), 0.f); std::fill(t->v.begin(), t->v.end(), 0.f);
  t->betap[0] = t->b1; t->betap[1] = t->b2;
}

// gradient(ps) do lossTot(net, x, y) end on one batch.  loss_out = total, policy, value, feature.  want_grad = 0: loss only.
int orc_trainer_loss_grad(void* h, const int8_t* state, const float* ypol, const float* yval, const int8_t* fstate, int64_t B64, float* loss_out,
                          int want_grad) {
  Trainer* t = (Trainer*)h;
  const int B = (int)B64, in = t->in, n = t->n, k = t->k, A = t->A, FS = t->FS, NH = t->NH;
  const float* W0 = t->params.data();
  const float* Wh = t->params.data() + t->off_heads;
  const float* bh = t->params.data() + t->off_bias;
  std::vector<float> X((size_t)B * in), H((size_t)(k + 1) * B * n), R((size_t)(k > 0 ? k : 1) * B * n), Zh((size_t)B * NH), dZh((size_t)B * NH);
  for (size_t i = 0; i < X.size(); i++) X[i] = (float)state[i];
  const size_t bn = (size_t)B * n;
  // ---- forward ----
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; b++) {
    const float* x = &X[(size_t)b * in];
    float* h0 = &H[(size_t)b * n];
    for (int o = 0; o < n; o++) h0[o] = relu(dot([&](int i) { return x[i]; }, [&](int i) { return W0[o + (size_t)n * i]; }, 0, in));
    for (int l = 1; l <= k; l++) {
      const float* W = t->params.data() + t->off_res + (size_t)(l - 1) * n * n;
      const float* hp = &H[(l - 1) * bn + (size_t)b * n];
      float* hl = &H[l * bn + (size_t)b * n];
      float* rl = &R[(l - 1) * bn + (size_t)b * n];
      for (int o = 0; o < n; o++) {
        const float r = relu(dot([&](int i) { return hp[i]; }, [&](int i) { return W[o + (size_t)n * i]; }, 0, n));
        rl[o] = r;
        hl[o] = relu(hp[o] + r);
      }
    }
    const float* hk = &H[k * bn + (size_t)b * n];
    for (int o = 0; o < NH; o++) Zh[(size_t)b * NH + o] = dot([&](int i) { return hk[i]; }, [&](int i) { return Wh[o + (size_t)NH * i]; }, 0, n) + bh[o];
  }
  // ---- heads, loss, dZh ----
  std::vector<float> terms((size_t)B * 3);
  const float invB = 1.0f / (float)B;
  const float invBF = 1.0f / ((float)B * (float)FS);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; b++) {
    const float* z = &Zh[(size_t)b * NH];
    float* dz = &dZh[(size_t)b * NH];
    float m = z[0];
    for (int a = 1; a < A; a++) m = z[a] > m ? z[a] : m;
    float s = 0.f;
    for (int a = 0; a < A; a++) s = s + c_expf(z[a] - m);
    const float ls = logf(s);
    float lp = 0.f, sumy = 0.f;
    const float* y = ypol + (size_t)b * A;
    for (int a = 0; a < A; a++) {
      lp = lp - y[a] * ((z[a] - m) - ls);
      sumy = sumy + y[a];
    }
    const float sy = sumy * invB;
    for (int a = 0; a < A; a++) dz[a] = (c_expf(z[a] - m) / s) * sy - y[a] * invB;
    const float v = c_sigmoidf(z[A]);
    const float dv = v - yval[b];
    dz[A] = ((2.0f * dv) * invB) * (v * (1.0f - v));
    float lf = 0.f;
    for (int j = 0; j < FS; j++) {
      const float f = c_tanhf(z[A + 1 + j]);
      const float df = f - (float)fstate[(size_t)b * FS + j];
      lf = lf + df * df;
      dz[A + 1 + j] = (((t->fweight * 2.0f) * df) * invBF) * (1.0f - f * f);
    }
    terms[(size_t)b * 3 + 0] = lp; terms[(size_t)b * 3 + 1] = dv * dv; terms[(size_t)b * 3 + 2] = lf;
  }
  double sl[3] = {0, 0, 0};
  for (int b = 0; b < B; b++) for (int c = 0; c < 3; c++) sl[c] += (double)terms[(size_t)b * 3 + c];
  const double lp = sl[0] / B, lv = sl[1] / B, lf = sl[2] / ((double)B * FS);
  loss_out[0] = (float)(lp + lv + (double)t->fweight * lf); loss_out[1] = (float)lp; loss_out[2] = (float)lv; loss_out[3] = (float)lf;
  if (!want_grad) return 0;

  // ---- backward ----
  const int ns = (B + KSLICE - 1) / KSLICE;
  float* g = t->grads.data();
  // weight gradient of a layer: G(o,i) at o + outs*i = sum over slices (ascending) of the slice's fma chain over its samples
  auto wgrad = [&](float* G, int outs, int ins, const float* act, int lda, const float* dz, int ldz) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < ins; i++)
      for (int o = 0; o < outs; o++) {
        float tot = 0.f;
        for (int s = 0; s < ns; s++) {
          const int b0 = s * KSLICE, b1 = std::min(B, b0 + KSLICE);
          tot = tot + dot([&](int b) { return act[(size_t)b * lda + i]; }, [&](int b) { return dz[(size_t)b * ldz + o]; }, b0, b1);
        }
        G[o + (size_t)outs * i] = tot;
      }
  };
  wgrad(g + t->off_heads, NH, n, &H[k * bn], n, dZh.data(), NH);
  for (int o = 0; o < NH; o++) {
    float tot = 0.f;
    for (int s = 0; s < ns; s++) {
      const int b0 = s * KSLICE, b1 = std::min(B, b0 + KSLICE);
      float part = 0.f;
      for (int b = b0; b < b1; b++) part = part + dZh[(size_t)b * NH + o];
      tot = tot + part;
    }
    g[t->off_bias + o] = tot;
  }
  std::vector<float> dS(bn), dZ(bn), dS2(bn), dZ2(bn);
  // dH_k = dZh Wh, masks of block k (or of the base layer)
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; b++)
    for (int i = 0; i < n; i++) {
      const float* dz = &dZh[(size_t)b * NH];
      const float dh = dot([&](int o) { return dz[o]; }, [&](int o) { return Wh[o + (size_t)NH * i]; }, 0, NH);
      const size_t idx = (size_t)b * n + i;
      const float ds = H[k * bn + idx] > 0.f ? dh : 0.f;
      if (k > 0) { dS[idx] = ds; dZ[idx] = R[(k - 1) * bn + idx] > 0.f ? ds : 0.f; } else { dZ[idx] = ds; }
    }
  for (int l = k; l >= 1; l--) {
    const float* W = t->params.data() + t->off_res + (size_t)(l - 1) * n * n;
    wgrad(g + t->off_res + (size_t)(l - 1) * n * n, n, n, &H[(l - 1) * bn], n, dZ.data(), n);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++)
      for (int i = 0; i < n; i++) {
        const float* dz = &dZ[(size_t)b * n];
        const size_t idx = (size_t)b * n + i;
        const float dh = dot([&](int o) { return dz[o]; }, [&](int o) { return W[o + (size_t)n * i]; }, 0, n) + dS[idx];
        const float ds = H[(l - 1) * bn + idx] > 0.f ? dh : 0.f;
        if (l - 1 >= 1) { dS2[idx] = ds; dZ2[idx] = R[(l - 2) * bn + idx] > 0.f ? ds : 0.f; } else { dZ2[idx] = ds; }
      }
    dS.swap(dS2); dZ.swap(dZ2);
  }
  wgrad(g, n, in, X.data(), in, dZ.data(), n);
  return 0;
}

// Flux.update!(opt, ps, gs), opt = Optimiser(ADAM(eta,(b1,b2)), WeightDecay(wd)); gs scaled by gscale first
int orc_trainer_apply(void* h, float gscale) {
  Trainer* t = (Trainer*)h;
  const double b1 = t->b1, b2 = t->b2, bp1 = t->betap[0], bp2 = t->betap[1];
  for (int64_t p = 0; p < t->P; p++) {
    const float d = t->grads[p] * gscale;
    const float m = (float)(b1 * (double)t->m[p] + (1.0 - b1) * (double)d);
    const float v = (float)(b2 * (double)t->v[p] + (1.0 - b2) * (double)(d * d));
    t->m[p] = m; t->v[p] = v;
    float delta = (float)((double)m / (1.0 - bp1) / (std::sqrt((double)v / (1.0 - bp2)) + t->eps) * t->lr);
    delta = (float)((double)delta + t->wd * (double)t->params[p]);
    t->params[p] = t->params[p] - delta;
  }
  t->betap[0] *= b1; t->betap[1] *= b2;
  return 0;
}

}  // extern "C"
