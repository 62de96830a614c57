// COMPILE-ONLY STAND-IN for jaxlib's xla/ffi/api/ffi.h -- NOT the real header, never shipped or linked.
//
// jax / jaxlib are not installed in this image, so the XLA-FFI shim (../../../../xla_ffi_shim.cc) cannot be built
// against the real API here.  This file declares just the names the shim uses -- data types, Error, AnyBuffer /
// Buffer<dtype>, Result<>, ScratchAllocator, PlatformStream<>, RemainingArgs / RemainingRets, the Ffi::Bind()
// builder and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with enough typing that `make xla-ffi-check` (g++ -fsyntax-only)
// catches signature rot: every handler's parameter list is static_assert-ed against its binding, and every
// fgx_* launcher call is checked against include/foragax_b200.h.  The spellings follow the public XLA FFI API as
// of jax 0.4.38+ (from memory: verify against a real jaxlib before relying on the shim).
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
enum class ErrorCode { kOk, kInvalidArgument, kInternal, kResourceExhausted, kUnimplemented };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : v_(std::move(v)) {}
  ErrorOr(Error) {}
  bool has_value() const { return v_.has_value(); }
  T& operator*() { return *v_; }
  T* operator->() { return &*v_; }

 private:
  std::optional<T> v_;
};

template <DataType dt> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<U32> { using type = uint32_t; };
template <> struct NativeTypeOf<U8> { using type = uint8_t; };
template <> struct NativeTypeOf<PRED> { using type = bool; };

class AnyBuffer {
 public:
  DataType element_type() const { return F32; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
  void* untyped_data() const { return nullptr; }
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeTypeOf<dt>::type;
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; }
};

template <typename T> struct PlatformStream {};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<T> get(size_t) const { return Error(); }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<Result<T>> get(size_t) const { return Error(); }
};

namespace internal {
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., typename CtxArg<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingRets> RemainingRets() const { return {}; }
  template <typename Fn>
  void To(Fn&&) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "XLA FFI handler: parameter list does not match its binding");
  }
};
}  // namespace internal

struct Ffi {
  static internal::Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)        \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame* call_frame) { \
    (binding).To(impl);                                           \
    (void)call_frame;                                             \
    return nullptr;                                               \
  }
