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

// ---- gradients: the two halves of jax.custom_vjp ----------------------
// forward rule: same signature as VitB200Forward, but keeps the activations inside the model
static ffi::Error VitB200TrainForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> images,
                                          ffi::ResultBuffer<ffi::F32> logits, int64_t handle) {
  auto* model = reinterpret_cast<vitb200_model*>(handle);
  if (model == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: null model handle");
  const auto dims = images.dimensions();
  if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: images must be rank 4");
  if (vitb200_train_forward(model, stream, images.typed_data(), static_cast<int>(dims[0]), logits->typed_data()) != 0)
    return ffi::Error(ffi::ErrorCode::kInternal, vitb200_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitB200TrainForward, VitB200TrainForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("handle"));

// backward rule: cotangent of the logits in, ONE flat f32 buffer of all leaf gradients out (the layout of
// vitb200_grads_buffer: leaves in vitb200_param_info order, each padded to a multiple of 64 floats); the
// Python side slices and reshapes it back into the params pytree.
static ffi::Error VitB200BackwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> dlogits,
                                      ffi::ResultBuffer<ffi::F32> grads_flat, int64_t handle) {
  auto* model = reinterpret_cast<vitb200_model*>(handle);
  if (model == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: null model handle");
  const auto dims = dlogits.dimensions();
  if (dims.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: dlogits must be rank 2");
  if (vitb200_backward(model, stream, dlogits.typed_data(), static_cast<int>(dims[0])) != 0)
    return ffi::Error(ffi::ErrorCode::kInternal, vitb200_last_error());
  float* src = nullptr;
  int64_t count = 0;
  if (vitb200_grads_buffer(model, &src, &count) != 0) return ffi::Error(ffi::ErrorCode::kInternal, vitb200_last_error());
  if (static_cast<int64_t>(grads_flat->element_count()) != count)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "vitb200: grads_flat has the wrong length");
  if (cudaMemcpyAsync(grads_flat->typed_data(), src, static_cast<size_t>(count) * sizeof(float),
                      cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "vitb200: copying the gradients out failed");
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitB200Backward, VitB200BackwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("handle"));
