// TensorFlow custom ops over the C ABI of libeodm_b200.so -- the layer
// tf.load_op_library("libeodm_tf.so") loads into the reference's TF2 model.
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: TensorFlow (headers and runtime) is not
// installable there, so this file has never been compiled; it is the thin glue a
// maintainer builds where TF lives (see INTEGRATION.md for the command).  All
// arithmetic is behind the C ABI, which is what the parity tests exercise.
//
//   EodmCounts(px f32[B,T,V], mask bool[B,T], kernel f32[n,V,K] (host)) -> S f32[K], N f32[]
//   EodmCountsGrad(px, mask, kernel, gS f32[K])                         -> dpx f32[B,T,V]
//   EodmNgramProb(px, kernel)                                           -> p f32[B,T-n+1,K]
//   EodmNgramProbGrad(px, kernel, dp)                                   -> dpx
// GPU kernels only: there is no CPU registration and no fallback.
#include <cstdint>
#include <mutex>
#include <unordered_map>

#include "../../include/eodm_b200.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#define EIGEN_USE_GPU
#include "unsupported/Eigen/CXX11/Tensor"

using namespace tensorflow;

namespace {

// One compact device table per distinct dense kernel buffer (the kernel is a frozen constant of the model:
// models/EODM.py:64-70 builds it once per run).
struct TableCache {
  std::mutex mu;
  std::unordered_map<const void*, eodm_table*> by_ptr;
  eodm_table* get(OpKernelContext* ctx, const Tensor& kernel, int device) {
    std::lock_guard<std::mutex> l(mu);
    const void* key = kernel.tensor_data().data();
    auto it = by_ptr.find(key);
    if (it != by_ptr.end()) return it->second;
    eodm_table* t = nullptr;
    int rc = eodm_table_create_from_dense(kernel.flat<float>().data(), (int)kernel.dim_size(0),
                                          (int)kernel.dim_size(1), (int)kernel.dim_size(2), device, &t);
    if (rc != EODM_OK) {
      ctx->SetStatus(errors::InvalidArgument("eodm_table_create_from_dense: ", eodm_last_error()));
      return nullptr;
    }
    by_ptr[key] = t;
    return t;
  }
};
TableCache g_tables;

void* gpu_stream(OpKernelContext* ctx) {
  return (void*)ctx->eigen_device<Eigen::GpuDevice>().stream();
}
int gpu_ordinal(OpKernelContext* ctx) {
  return ctx->op_device_context()->stream()->parent()->device_ordinal();
}
Status to_status(int rc) {
  if (rc == EODM_OK) return Status();
  return errors::InvalidArgument(eodm_last_error());   // -> tf.errors.InvalidArgumentError in Python
}

class EodmCountsOp : public OpKernel {
 public:
  explicit EodmCountsOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &mask = ctx->input(1), &kernel = ctx->input(2);
    OP_REQUIRES(ctx, px.dims() == 3 && mask.dims() == 2 && kernel.dims() == 3,
                errors::InvalidArgument("EodmCounts: px [B,T,V], mask [B,T], kernel [n,V,K] expected"));
    eodm_table* t = g_tables.get(ctx, kernel, gpu_ordinal(ctx));
    if (!t) return;
    const int B = px.dim_size(0), T = px.dim_size(1), K = kernel.dim_size(2);
    Tensor *S = nullptr, *N = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({K}), &S));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({}), &N));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64_t)eodm_workspace_bytes(t, B, T)}), &ws));
    OP_REQUIRES_OK(ctx, to_status(eodm_counts_fwd(t, px.flat<float>().data(), (const uint8_t*)mask.flat<bool>().data(),
                                                  B, T, S->flat<float>().data(), N->flat<float>().data(),
                                                  ws.flat<uint8>().data(), gpu_stream(ctx))));
  }
};

class EodmCountsGradOp : public OpKernel {
 public:
  explicit EodmCountsGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &mask = ctx->input(1), &kernel = ctx->input(2), &gS = ctx->input(3);
    eodm_table* t = g_tables.get(ctx, kernel, gpu_ordinal(ctx));
    if (!t) return;
    const int B = px.dim_size(0), T = px.dim_size(1);
    OP_REQUIRES(ctx, gS.NumElements() == kernel.dim_size(2), errors::InvalidArgument("EodmCountsGrad: len(gS) != K"));
    Tensor* dpx = nullptr;
    Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, px.shape(), &dpx));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64_t)eodm_workspace_bytes(t, B, T)}), &ws));
    OP_REQUIRES_OK(ctx, to_status(eodm_counts_bwd(t, px.flat<float>().data(), (const uint8_t*)mask.flat<bool>().data(),
                                                  B, T, gS.flat<float>().data(), dpx->flat<float>().data(),
                                                  ws.flat<uint8>().data(), gpu_stream(ctx))));
  }
};

class EodmNgramProbOp : public OpKernel {
 public:
  explicit EodmNgramProbOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &kernel = ctx->input(1);
    eodm_table* t = g_tables.get(ctx, kernel, gpu_ordinal(ctx));
    if (!t) return;
    const int B = px.dim_size(0), T = px.dim_size(1), n = kernel.dim_size(0), K = kernel.dim_size(2);
    OP_REQUIRES(ctx, T >= n, errors::InvalidArgument("EodmNgramProb: T < kernel_size, Conv1D 'valid' has no output"));
    Tensor* p = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, T - n + 1, K}), &p));
    OP_REQUIRES_OK(ctx, to_status(eodm_prob_fwd(t, px.flat<float>().data(), B, T, p->flat<float>().data(),
                                                gpu_stream(ctx))));
  }
};

class EodmNgramProbGradOp : public OpKernel {
 public:
  explicit EodmNgramProbGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &kernel = ctx->input(1), &dp = ctx->input(2);
    eodm_table* t = g_tables.get(ctx, kernel, gpu_ordinal(ctx));
    if (!t) return;
    Tensor* dpx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, px.shape(), &dpx));
    OP_REQUIRES_OK(ctx, to_status(eodm_prob_bwd(t, px.flat<float>().data(), dp.flat<float>().data(),
                                                (int)px.dim_size(0), (int)px.dim_size(1), dpx->flat<float>().data(),
                                                gpu_stream(ctx))));
  }
};

}  // namespace

REGISTER_OP("EodmCounts").Input("px: float").Input("mask: bool").Input("kernel: float")
    .Output("s: float").Output("n: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle k;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 3, &k));
      c->set_output(0, c->Vector(c->Dim(k, 2)));
      c->set_output(1, c->Scalar());
      return Status();
    });
REGISTER_OP("EodmCountsGrad").Input("px: float").Input("mask: bool").Input("kernel: float").Input("gs: float")
    .Output("dpx: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return Status(); });
REGISTER_OP("EodmNgramProb").Input("px: float").Input("kernel: float").Output("p: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle x, k;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &x));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &k));
      c->set_output(0, c->MakeShape({c->Dim(x, 0), c->UnknownDim(), c->Dim(k, 2)}));
      return Status();
    });
REGISTER_OP("EodmNgramProbGrad").Input("px: float").Input("kernel: float").Input("dp: float").Output("dpx: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return Status(); });

// GPU only; `kernel` stays in host memory (it is compacted on the host once and cached).
REGISTER_KERNEL_BUILDER(Name("EodmCounts").Device(DEVICE_GPU).HostMemory("kernel"), EodmCountsOp);
REGISTER_KERNEL_BUILDER(Name("EodmCountsGrad").Device(DEVICE_GPU).HostMemory("kernel"), EodmCountsGradOp);
REGISTER_KERNEL_BUILDER(Name("EodmNgramProb").Device(DEVICE_GPU).HostMemory("kernel"), EodmNgramProbOp);
REGISTER_KERNEL_BUILDER(Name("EodmNgramProbGrad").Device(DEVICE_GPU).HostMemory("kernel"), EodmNgramProbGradOp);
