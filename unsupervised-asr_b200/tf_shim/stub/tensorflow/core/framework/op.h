// MINIMAL STAND-IN (compile-only check, see op_kernel.h in this directory): REGISTER_OP and its builder chain.
#ifndef EODM_TF_STUB_OP_H_
#define EODM_TF_STUB_OP_H_
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
namespace tensorflow {
struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*);
  OpDefBuilderStub& Output(const char*);
  OpDefBuilderStub& Attr(const char*);
  OpDefBuilderStub& SetShapeFn(Status (*fn)(shape_inference::InferenceContext*));
};
OpDefBuilderStub RegisterOpStub(const char* name);
#define REGISTER_OP(NAME) \
  static ::tensorflow::OpDefBuilderStub& EODM_STUB_CAT(eodm_stub_op_, __LINE__) __attribute__((unused)) = ::tensorflow::RegisterOpStub(NAME)
}  // namespace tensorflow
#endif
