// oracle/ref_shim: stand-in for <torch/serialize/tensor.h> used ONLY to compile the
// reference's *_gpu.cu files (which need nothing from torch except the name at::Tensor in
// wrapper prototypes they never call).  Declaring a function that takes an incomplete class
// type by value is legal C++, so a forward declaration is all the reference headers need.
#pragma once
namespace at { class Tensor; }
