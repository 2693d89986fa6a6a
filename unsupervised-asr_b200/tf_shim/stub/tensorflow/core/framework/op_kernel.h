// MINIMAL STAND-IN for the TensorFlow C++ op API, for a COMPILE-ONLY check of tf_shim/eodm_tf_ops.cc in an image where
// TensorFlow cannot be installed (`make -C tf_shim check`).  It declares exactly the names that file uses, with the
// argument and return types of TensorFlow 2.x, and implements nothing: the object file it yields is never linked or
// loaded.  A real build uses the headers of the installed TensorFlow instead.
#ifndef EODM_TF_STUB_OP_KERNEL_H_
#define EODM_TF_STUB_OP_KERNEL_H_
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

namespace Eigen {
struct GpuDevice {
  void* stream() const;   // cudaStream_t in the real header
};
}  // namespace Eigen

namespace tensorflow {
typedef unsigned char uint8;
typedef long long int64;
enum DataType { DT_FLOAT = 1, DT_UINT8 = 4, DT_BOOL = 10 };

class Status {
 public:
  Status() {}
  bool ok() const;
};
namespace errors {
template <typename... Args>
Status InvalidArgument(Args... args);
}

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t> dims);
  bool operator==(const TensorShape& o) const;
};
struct StringPiece {
  const char* data() const;
  size_t size() const;
};
template <typename T>
struct Flat {
  T* data() const;
};
class Tensor {
 public:
  Tensor();
  int dims() const;
  int64_t dim_size(int d) const;
  int64_t NumElements() const;
  const TensorShape& shape() const;
  StringPiece tensor_data() const;
  template <typename T> Flat<T> flat();
  template <typename T> Flat<const T> flat() const;
};

namespace se_stub {
struct Executor { int device_ordinal() const; };
struct Stream { Executor* parent() const; };
}  // namespace se_stub
class DeviceContext {
 public:
  se_stub::Stream* stream() const;
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const;
  void SetStatus(const Status& s);
};
class OpKernelContext {
 public:
  const Tensor& input(int i);
  Status allocate_output(int i, const TensorShape& shape, Tensor** out);
  Status allocate_temp(DataType t, const TensorShape& shape, Tensor* out);
  void SetStatus(const Status& s);
  template <typename D> const D& eigen_device() const;
  DeviceContext* op_device_context();
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* ctx) = 0;
};

#define OP_REQUIRES(CTX, COND, STATUS) \
  do {                                 \
    if (!(COND)) {                     \
      (CTX)->SetStatus(STATUS);        \
      return;                          \
    }                                  \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)           \
  do {                                     \
    ::tensorflow::Status s_(__VA_ARGS__);  \
    if (!s_.ok()) {                        \
      (CTX)->SetStatus(s_);                \
      return;                              \
    }                                      \
  } while (0)
#define TF_RETURN_IF_ERROR(...)            \
  do {                                     \
    ::tensorflow::Status s_(__VA_ARGS__);  \
    if (!s_.ok()) return s_;               \
  } while (0)

constexpr const char* DEVICE_GPU = "GPU";
struct KernelDefBuilder {
  KernelDefBuilder& Device(const char*);
  KernelDefBuilder& HostMemory(const char*);
};
KernelDefBuilder Name(const char*);
#define EODM_STUB_CAT2(a, b) a##b
#define EODM_STUB_CAT(a, b) EODM_STUB_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(BUILDER, CLS) \
  static ::tensorflow::OpKernel* EODM_STUB_CAT(eodm_stub_make_, __LINE__)(::tensorflow::OpKernelConstruction* c) { return new CLS(c); }
}  // namespace tensorflow
#endif
