// MINIMAL STAND-IN (compile-only check, see op_kernel.h in this directory): the shape-inference names the shim uses.
#ifndef EODM_TF_STUB_SHAPE_INFERENCE_H_
#define EODM_TF_STUB_SHAPE_INFERENCE_H_
#include <initializer_list>
#include "tensorflow/core/framework/op_kernel.h"
namespace tensorflow {
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int i);
  Status WithRank(ShapeHandle s, int rank, ShapeHandle* out);
  DimensionHandle Dim(ShapeHandle s, int i);
  DimensionHandle UnknownDim();
  ShapeHandle Vector(DimensionHandle d);
  ShapeHandle Scalar();
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle> dims);
  void set_output(int i, ShapeHandle s);
};
}  // namespace shape_inference
}  // namespace tensorflow
#endif
