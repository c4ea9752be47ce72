// vitb200_xla_ffi.cc -- XLA FFI (jax.ffi) adaptor over the C ABI of include/vitb200.h.
//
// NOT built in this repository's image: it needs the XLA FFI headers that ship inside jaxlib
// (jaxlib/include/xla/ffi/api/{c_api,api,ffi}.h) and JAX/jaxlib cannot be installed here
// (INTEGRATION.md section B).  On a machine that has jaxlib:
//
//   INC=$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))")
//   g++ -O2 -std=c++17 -shared -fPIC -I$INC -I../../include vitb200_xla_ffi.cc \
//       -L../../vit_flax_b200 -lvitb200 -Wl,-rpath,'$ORIGIN/../../vit_flax_b200' -o libvitb200_xla.so
//
// One handler: images f32[B,H,W,C] on the device in, logits f32[B,num_classes] out, on the stream
// XLA hands over; the model handle (created and loaded through the C ABI, see jax_binding.py) is an
// int64 attribute.  vitb200_forward only enqueues kernels: no synchronisation inside the call.
#include <cstdint>

#include "vitb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error VitB200ForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> images,
                                     ffi::ResultBuffer<ffi::F32> logits, int64_t handle) {
  auto* model = reinterpret_cast<vitb200_model*>(handle);
  if (model == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: null model handle");
  const auto dims = images.dimensions();
  if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: images must be rank 4");
  const int batch = static_cast<int>(dims[0]);
  if (vitb200_forward(model, stream, images.typed_data(), batch, logits->typed_data()) != 0)
    return ffi::Error(ffi::ErrorCode::kInternal, vitb200_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitB200Forward, VitB200ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("handle"));
