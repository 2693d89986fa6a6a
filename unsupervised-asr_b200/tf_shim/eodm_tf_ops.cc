// TensorFlow custom ops over the C ABI of libeodm_b200.so -- the layer
// tf.load_op_library("libeodm_tf.so") loads into the reference's TF2 model.
//
// TensorFlow (headers and runtime) is not installable in this repository's image, so this file is only ever COMPILED
// here, against the minimal stand-in headers of tf_shim/stub/ (`make -C tf_shim check`, run by __graft_entry__.build());
// a maintainer builds the real libeodm_tf.so where TensorFlow lives.  All arithmetic is behind the C
// ABI, which is what the parity tests exercise.
//
//   EodmCounts(px f32[B,T,V], mask bool[B,T], kernel f32[n,V,K] (host); table_id) -> S f32[K], N f32[]
//   EodmCountsGrad(px, mask, kernel, gS f32[K]; table_id)                         -> dpx f32[B,T,V]
//   EodmNgramProb(px, kernel; table_id)                                           -> p f32[B,T-n+1,K]
//   EodmNgramProbGrad(px, kernel, dp; table_id)                                   -> dpx
//   EodmLoss(logits f32[B,T,V], mask bool[B,T], kernel (host), py f32[K] (host); table_id) -> loss f32[], dlogits f32[B,T,V]
//       the fused step of models/EODM.py:5-25 + its gradient (eodm_session_step_device): softmax, counts, loss,
//       counts VJP and softmax VJP in one op -- what bench.py times
// GPU kernels only: there is no CPU registration and no fallback.
//
// Table identity.  `kernel` is a frozen constant of the model (models/EODM.py:64-70 builds it once per run), but with
// HostMemory("kernel") TensorFlow may hand the op a fresh host copy on every execution, so the host ADDRESS says nothing.
// Every P_Ngram instance therefore carries a process-unique `table_id` attribute (assigned in tf_shim/models/EODM.py);
// the compact device table is built once per (table_id, GPU ordinal) from the kernel's contents, checked against the
// kernel's shape on every later use, and freed when the library is unloaded.
#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/eodm_b200.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#define EIGEN_USE_GPU
#include "unsupported/Eigen/CXX11/Tensor"

using namespace tensorflow;

namespace {

struct TableEntry {
  eodm_table* table = nullptr;
  int n = 0, V = 0, K = 0;
  eodm_session* session = nullptr;   // fused step (EodmLoss); rebuilt when a larger batch arrives
  int maxB = 0, maxT = 0;
};

class TableRegistry {
 public:
  ~TableRegistry() {
    for (auto& kv : entries_) {
      if (kv.second.session) eodm_session_destroy(kv.second.session);
      if (kv.second.table) eodm_table_destroy(kv.second.table);
    }
  }
  // the table of (table_id, device), built from `kernel` on first use; nullptr + ctx status on error
  TableEntry* get(OpKernelContext* ctx, int64_t table_id, const Tensor& kernel, int device) {
    std::lock_guard<std::mutex> l(mu_);
    if (kernel.dims() != 3) {
      ctx->SetStatus(errors::InvalidArgument("kernel must be [n, V, K]"));
      return nullptr;
    }
    const int n = (int)kernel.dim_size(0), V = (int)kernel.dim_size(1), K = (int)kernel.dim_size(2);
    TableEntry& e = entries_[std::make_pair(table_id, device)];
    if (e.table) {
      if (e.n != n || e.V != V || e.K != K) {
        ctx->SetStatus(errors::InvalidArgument("table_id ", table_id, " was built for a kernel of another shape"));
        return nullptr;
      }
      return &e;
    }
    const int rc = eodm_table_create_from_dense(kernel.flat<float>().data(), n, V, K, device, &e.table);
    if (rc != EODM_OK) {
      ctx->SetStatus(errors::InvalidArgument("eodm_table_create_from_dense: ", eodm_last_error()));
      e.table = nullptr;
      return nullptr;
    }
    e.n = n; e.V = V; e.K = K;
    return &e;
  }
  // the fused-step session of an entry, large enough for [B, T]
  eodm_session* session(OpKernelContext* ctx, TableEntry* e, const Tensor& py, int B, int T) {
    std::lock_guard<std::mutex> l(mu_);
    if (e->session && B <= e->maxB && T <= e->maxT) return e->session;
    if (e->session) eodm_session_destroy(e->session);
    e->session = nullptr;
    const int mb = B > e->maxB ? B : e->maxB, mt = T > e->maxT ? T : e->maxT;
    const int rc = eodm_session_create(e->table, py.flat<float>().data(), mb, mt, &e->session);
    if (rc != EODM_OK) {
      ctx->SetStatus(errors::InvalidArgument("eodm_session_create: ", eodm_last_error()));
      return nullptr;
    }
    e->maxB = mb; e->maxT = mt;
    return e->session;
  }

 private:
  std::mutex mu_;
  std::map<std::pair<int64_t, int>, TableEntry> entries_;
};
TableRegistry g_tables;

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

class TableOp : public OpKernel {
 public:
  explicit TableOp(OpKernelConstruction* c) : OpKernel(c) {
    Status s = c->GetAttr("table_id", &table_id_);
    if (!s.ok()) c->SetStatus(s);
  }

 protected:
  int64_t table_id_ = 0;
};

class EodmCountsOp : public TableOp {
 public:
  explicit EodmCountsOp(OpKernelConstruction* c) : TableOp(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &mask = ctx->input(1), &kernel = ctx->input(2);
    OP_REQUIRES(ctx, px.dims() == 3 && mask.dims() == 2 && kernel.dims() == 3,
                errors::InvalidArgument("EodmCounts: px [B,T,V], mask [B,T], kernel [n,V,K] expected"));
    TableEntry* e = g_tables.get(ctx, table_id_, kernel, gpu_ordinal(ctx));
    if (!e) return;
    const int B = px.dim_size(0), T = px.dim_size(1), K = e->K;
    Tensor *S = nullptr, *N = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({K}), &S));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({}), &N));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64_t)eodm_workspace_bytes(e->table, B, T)}), &ws));
    OP_REQUIRES_OK(ctx, to_status(eodm_counts_fwd(e->table, px.flat<float>().data(),
                                                  (const uint8_t*)mask.flat<bool>().data(), B, T, S->flat<float>().data(),
                                                  N->flat<float>().data(), ws.flat<uint8>().data(), gpu_stream(ctx))));
  }
};

class EodmCountsGradOp : public TableOp {
 public:
  explicit EodmCountsGradOp(OpKernelConstruction* c) : TableOp(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &mask = ctx->input(1), &kernel = ctx->input(2), &gS = ctx->input(3);
    TableEntry* e = g_tables.get(ctx, table_id_, kernel, gpu_ordinal(ctx));
    if (!e) return;
    const int B = px.dim_size(0), T = px.dim_size(1);
    OP_REQUIRES(ctx, gS.NumElements() == e->K, errors::InvalidArgument("EodmCountsGrad: len(gS) != K"));
    Tensor* dpx = nullptr;
    Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, px.shape(), &dpx));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64_t)eodm_workspace_bytes(e->table, B, T)}), &ws));
    OP_REQUIRES_OK(ctx, to_status(eodm_counts_bwd(e->table, px.flat<float>().data(),
                                                  (const uint8_t*)mask.flat<bool>().data(), B, T, gS.flat<float>().data(),
                                                  dpx->flat<float>().data(), ws.flat<uint8>().data(), gpu_stream(ctx))));
  }
};

class EodmNgramProbOp : public TableOp {
 public:
  explicit EodmNgramProbOp(OpKernelConstruction* c) : TableOp(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &kernel = ctx->input(1);
    TableEntry* e = g_tables.get(ctx, table_id_, kernel, gpu_ordinal(ctx));
    if (!e) return;
    const int B = px.dim_size(0), T = px.dim_size(1);
    OP_REQUIRES(ctx, T >= e->n, errors::InvalidArgument("EodmNgramProb: T < kernel_size, Conv1D 'valid' has no output"));
    Tensor* p = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, T - e->n + 1, e->K}), &p));
    OP_REQUIRES_OK(ctx, to_status(eodm_prob_fwd(e->table, px.flat<float>().data(), B, T, p->flat<float>().data(),
                                                gpu_stream(ctx))));
  }
};

class EodmNgramProbGradOp : public TableOp {
 public:
  explicit EodmNgramProbGradOp(OpKernelConstruction* c) : TableOp(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &px = ctx->input(0), &kernel = ctx->input(1), &dp = ctx->input(2);
    TableEntry* e = g_tables.get(ctx, table_id_, kernel, gpu_ordinal(ctx));
    if (!e) return;
    Tensor* dpx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, px.shape(), &dpx));
    OP_REQUIRES_OK(ctx, to_status(eodm_prob_bwd(e->table, px.flat<float>().data(), dp.flat<float>().data(),
                                                (int)px.dim_size(0), (int)px.dim_size(1), dpx->flat<float>().data(),
                                                gpu_stream(ctx))));
  }
};

// The fused step: EODM_loss (models/EODM.py:5-25) and its gradient wrt `_logits` (the tape of main_EODM.py:168) in
// one op on the op's stream; the session owns the scratch (px, dpx, counts, workspace) between steps.
class EodmLossOp : public TableOp {
 public:
  explicit EodmLossOp(OpKernelConstruction* c) : TableOp(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &logits = ctx->input(0), &mask = ctx->input(1), &kernel = ctx->input(2), &py = ctx->input(3);
    OP_REQUIRES(ctx, logits.dims() == 3 && mask.dims() == 2,
                errors::InvalidArgument("EodmLoss: logits [B,T,V] and mask [B,T] expected"));
    TableEntry* e = g_tables.get(ctx, table_id_, kernel, gpu_ordinal(ctx));
    if (!e) return;
    OP_REQUIRES(ctx, py.NumElements() == e->K, errors::InvalidArgument("EodmLoss: len(py) != K"));
    OP_REQUIRES(ctx, logits.dim_size(2) == e->V, errors::InvalidArgument("EodmLoss: logits have another vocabulary size"));
    const int B = logits.dim_size(0), T = logits.dim_size(1);
    eodm_session* s = g_tables.session(ctx, e, py, B, T);
    if (!s) return;
    Tensor *loss = nullptr, *dlogits = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, logits.shape(), &dlogits));
    OP_REQUIRES_OK(ctx, to_status(eodm_session_step_device(s, logits.flat<float>().data(),
                                                           (const uint8_t*)mask.flat<bool>().data(), B, T, nullptr,
                                                           loss->flat<float>().data(), dlogits->flat<float>().data(),
                                                           gpu_stream(ctx))));
  }
};

}  // namespace

REGISTER_OP("EodmCounts").Input("px: float").Input("mask: bool").Input("kernel: float").Attr("table_id: int")
    .Output("s: float").Output("n: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle k;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 3, &k));
      c->set_output(0, c->Vector(c->Dim(k, 2)));
      c->set_output(1, c->Scalar());
      return Status();
    });
REGISTER_OP("EodmCountsGrad").Input("px: float").Input("mask: bool").Input("kernel: float").Input("gs: float")
    .Attr("table_id: int").Output("dpx: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return Status(); });
REGISTER_OP("EodmNgramProb").Input("px: float").Input("kernel: float").Attr("table_id: int").Output("p: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle x, k;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &x));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &k));
      c->set_output(0, c->MakeShape({c->Dim(x, 0), c->UnknownDim(), c->Dim(k, 2)}));
      return Status();
    });
REGISTER_OP("EodmNgramProbGrad").Input("px: float").Input("kernel: float").Input("dp: float").Attr("table_id: int")
    .Output("dpx: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return Status(); });
REGISTER_OP("EodmLoss").Input("logits: float").Input("mask: bool").Input("kernel: float").Input("py: float")
    .Attr("table_id: int").Output("loss: float").Output("dlogits: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->Scalar());
      c->set_output(1, c->input(0));
      return Status();
    });

// GPU only; `kernel` (and `py` of the fused op) stay in host memory: they are consumed on the host once per table.
REGISTER_KERNEL_BUILDER(Name("EodmCounts").Device(DEVICE_GPU).HostMemory("kernel"), EodmCountsOp);
REGISTER_KERNEL_BUILDER(Name("EodmCountsGrad").Device(DEVICE_GPU).HostMemory("kernel"), EodmCountsGradOp);
REGISTER_KERNEL_BUILDER(Name("EodmNgramProb").Device(DEVICE_GPU).HostMemory("kernel"), EodmNgramProbOp);
REGISTER_KERNEL_BUILDER(Name("EodmNgramProbGrad").Device(DEVICE_GPU).HostMemory("kernel"), EodmNgramProbGradOp);
REGISTER_KERNEL_BUILDER(Name("EodmLoss").Device(DEVICE_GPU).HostMemory("kernel").HostMemory("py"), EodmLossOp);
