// A DECLARATION-ONLY STAND-IN for jaxlib's "xla/ffi/api/ffi.h", written from the public documentation of the XLA
// typed-FFI C++ API (Ffi::Bind().Ctx<>().Arg<>().Attr<>().Ret<>(), Buffer<dtype>, Result<>, Error, ScratchAllocator,
// PlatformStream<>, XLA_FFI_DEFINE_HANDLER_SYMBOL).  jaxlib is not installable in this image, so this header exists
// for ONE purpose: `g++ -fsyntax-only -I integration/stub -I include integration/xla_ffi.cc` type-checks the shim —
//   * every drt_* call against the real include/differt_b200.h (argument count, order and types), and
//   * every handler's parameter list against the binding it is registered with (the static_assert below),
// which is what tests/test_abi.py::test_xla_ffi_shim_type_checks runs.  It registers nothing and cannot execute; the
// real header must be used to build the shim (`jax.ffi.include_dir()`), and where the two disagree the real one wins.
#pragma once

#include <complex>
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>

extern "C" {
struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;
}

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, BF16, F32, F64, C64, C128 };

namespace internal {
template <DataType>
struct Native;
template <> struct Native<PRED> { using type = bool; };
template <> struct Native<S32> { using type = std::int32_t; };
template <> struct Native<S64> { using type = std::int64_t; };
template <> struct Native<U8> { using type = std::uint8_t; };
template <> struct Native<U32> { using type = std::uint32_t; };
template <> struct Native<F32> { using type = float; };
template <> struct Native<F64> { using type = double; };
template <> struct Native<C64> { using type = std::complex<float>; };
}  // namespace internal

template <typename T>
class Span {
   public:
    const T &operator[](std::size_t i) const;
    std::size_t size() const;
    const T *begin() const;
    const T *end() const;
};

template <DataType dtype>
class Buffer {
   public:
    using Native = typename internal::Native<dtype>::type;
    Native *typed_data() const;
    void *untyped_data() const;
    Span<std::int64_t> dimensions() const;
    std::size_t element_count() const;
    std::size_t size_bytes() const;
};

template <typename T>
class Result {
   public:
    T *operator->() const;
    T &operator*() const;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kNotFound, kResourceExhausted, kInternal, kUnimplemented };

class Error {
   public:
    Error();
    Error(ErrorCode code, std::string message);
    static Error Success();
    bool success() const;
};

class ScratchAllocator {
   public:
    std::optional<void *> Allocate(std::size_t size, std::size_t alignment = 1);
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename C>
struct CtxType { using type = C; };
template <typename T>
struct CtxType<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
    template <typename C> Binding<Ts..., typename CtxType<C>::type> Ctx() const;
    template <typename A> Binding<Ts..., A> Arg() const;
    template <typename A> Binding<Ts..., A> Attr(const char *name) const;
    template <typename R> Binding<Ts..., Result<R>> Ret() const;
    using Signature = Error (*)(Ts...);
};
}  // namespace internal

struct Ffi {
    static internal::Binding<> Bind();
};

}  // namespace xla::ffi

// the real macro defines `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)`; the stand-in declares it and checks that
// the implementation takes exactly what the binding decodes, in that order
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                          \
    static_assert(std::is_same_v<typename std::remove_cv_t<std::remove_reference_t<decltype(binding)>>::Signature, \
                                 decltype(&impl)>,                                                                  \
                  #name ": the handler's parameters do not match its binding");                                     \
    extern "C" XLA_FFI_Error *name(XLA_FFI_CallFrame *call_frame)
