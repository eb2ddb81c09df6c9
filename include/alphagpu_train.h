/* alphagpu_train.h — C ABI of the training step: what a `ccall` from train.jl binds instead of
 * Flux/Zygote/CUDA.jl for
 *
 *     custom_train!((x,y)->lossTot(net,x,y), Flux.params(net), [(x,(p,r,f))], Optimiser(ADAM(lr), WeightDecay(1e-4)))
 *                                                                                   (train.jl:12-15,47-51,91-96,128-162)
 *
 * on a `networkf` (DenseNet.jl:161-198): base Dense(in,n,relu) without bias, k blocks  b = relu(b + relu(W b)),
 * heads policy Dense(n,A), value Dense(n,1,sigmoid), feature Dense(n,FS,tanh), all three with bias.
 *
 *   loss = logitcrossentropy(p, y_policy) + mse(v, y_value) + 0.001 * mse(f, y_feature)            (train.jl:12-15)
 *
 * Conventions are those of alphagpu.h: plain pointers and sizes, 0 or a negative agpu_status, nothing throws, no CPU
 * fallback.  Weight arrays are Julia column-major fp32 exactly as Flux holds them (`Dense.weight` is out x in).  Batch
 * arrays are the SoA sample format of agpu_samples (state int8 [B][in], policy f32 [B][A], value f32 [B], fstate int8
 * [B][FS]) — what traininPipe copies into tmpx/tmpy/tmpr/tmpf (train.jl:85-95) — host or device pointers.
 *
 * Data-parallel training (BASELINE config 4): every rank calls agpu_trainer_loss_grad on its shard of the batch, the host
 * all-reduces the flat gradient (agpu_trainer_grad_buffer, NCCL through torch.distributed in alphagpu_b200/train.py),
 * then every rank calls agpu_trainer_apply(1/world_size).
 *
 * Arithmetic is fp32 and specified operation by operation (DESIGN.md "training step"): every dot product is an fma
 * chain ascending in k, weight gradients are summed per 256-sample slice and the slices added in order, Adam runs in
 * fp64 per element as Flux 0.12.6 does with its Float64 hyper-parameters.  The CUDA path and the CPU oracle
 * (oracle/train_oracle.cpp) agree bit for bit on one GPU.
 */
#ifndef ALPHAGPU_TRAIN_H
#define ALPHAGPU_TRAIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agpu_trainer agpu_trainer;

typedef struct agpu_train_config {
  int32_t device;       /* CUDA ordinal */
  int32_t in;           /* 2 * VectorizedState */
  int32_t width;        /* n */
  int32_t blocks;       /* k */
  int32_t actions;      /* A = maxActions */
  int32_t fsize;        /* FS = FeatureSize */
  int32_t max_batch;    /* largest B of a step (main4IARow.jl:102-105: 8192) */
  int32_t reserved;
  double lr;            /* ADAM eta  (train.jl:47: 0.001) */
  double beta1, beta2;  /* 0.9, 0.999 (Flux defaults) */
  double eps;           /* 1e-8 */
  double weight_decay;  /* WeightDecay(0.0001) (train.jl:50) */
  float feature_weight; /* 0.001f0 (train.jl:14) */
  float reserved2;
} agpu_train_config;      /* 80 bytes */

/* trainingnet = ressimplesf(in, A, FS, n, k) |> gpu  (main4IARow.jl:123): allocates parameters (zero), Adam state,
 * activations and gradient workspace for max_batch samples. */
int agpu_trainer_create(agpu_trainer** out, const agpu_train_config* cfg);
void agpu_trainer_destroy(agpu_trainer* tr);
const char* agpu_trainer_last_error(const agpu_trainer* tr);   /* tr == NULL: error of the last failed create */

/* Flux.params(net) in: base (n x in), res[k] (n x n), policy (A x n) + bias (A), value (1 x n) + bias (1), feature
 * (FS x n) + bias (FS).  reset_optimizer != 0 also forgets the Adam moments (a fresh `opt`, train.jl:50). */
int agpu_trainer_set_params(agpu_trainer* tr, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                            const float* val_w, const float* val_b, const float* feat_w, const float* feat_b, int32_t reset_optimizer);
/* to_cpu(net) / convert_back(net) (DenseNet.jl:170,331): the same arrays out (any pointer may be NULL = skip). */
int agpu_trainer_get_params(agpu_trainer* tr, float* base, float* const* res, float* pol_w, float* pol_b, float* val_w, float* val_b,
                            float* feat_w, float* feat_b);
/* the gradient of the last agpu_trainer_loss_grad in the same shapes (tests, diagnostics) */
int agpu_trainer_get_grads(agpu_trainer* tr, float* base, float* const* res, float* pol_w, float* pol_b, float* val_w, float* val_b,
                           float* feat_w, float* feat_b);

/* gradient(ps) do lossTot(net, x, y) end (train.jl:133-136) on one batch: forward, loss, backward into the flat gradient.
 * loss_out[4] = total, policy, value, feature (unweighted). */
int agpu_trainer_loss_grad(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate,
                           int64_t B, float loss_out[4]);
/* the flat fp32 gradient on the device (count elements; order: base, res[0..k), heads packed (A+1+FS) x n, head biases)
 * for the caller's all-reduce */
int agpu_trainer_grad_buffer(agpu_trainer* tr, void** device_ptr, int64_t* count);
/* Flux.update!(opt, ps, gs) (train.jl:158) with gs scaled by grad_scale (1/world_size after a sum all-reduce) */
int agpu_trainer_apply(agpu_trainer* tr, float grad_scale);
/* loss_grad + apply(1): one custom_train! iteration on one GPU */
int agpu_trainer_step(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B,
                      float loss_out[4]);
/* forward only: lossTot without the gradient (validation), same loss_out */
int agpu_trainer_loss(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B,
                      float loss_out[4]);

/* optimiser state for checkpoints: flat first / second moments (count of agpu_trainer_grad_buffer) and the running
 * beta powers; set != 0 installs, else reads. */
int agpu_trainer_opt_state(agpu_trainer* tr, float* m, float* v, double beta_pow[2], int32_t set);
/* device time of the last step's kernels in milliseconds (CUDA events on the trainer's stream): [0] loss_grad, [1] apply */
int agpu_trainer_last_ms(agpu_trainer* tr, float ms[2]);

#ifdef __cplusplus
}
#endif
#endif /* ALPHAGPU_TRAIN_H */
