// TEST INFRASTRUCTURE: a minimal stand-in for jaxlib's xla/ffi/api/ffi.h (which is not installable in this image) with the
// same names and shapes the shim uses -- Buffer / ResultBuffer / Span / Error / PlatformStream / Ffi::Bind() ... .To(fn) and
// XLA_FFI_DEFINE_HANDLER_SYMBOL -- so that chain_control_b200/csrc/xla_ffi_shim.cc is compiled (and its handler signatures
// are checked against their bindings by a static_assert) in the CPU test-suite and cannot rot.  Nothing here is shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

namespace xla {
namespace ffi {

enum class DataType { F32, S32, U8, F64 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType S32 = DataType::S32;

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  const T* data() const { return d_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_ = nullptr;
  size_t n_ = 0;
};

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  const std::string& message() const { return msg_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <DataType dtype>
class Buffer {
 public:
  using Native = std::conditional_t<dtype == DataType::F32, float, std::conditional_t<dtype == DataType::F64, double, std::conditional_t<dtype == DataType::S32, int32_t, uint8_t>>>;
  Native* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
    return n;
  }
  size_t size_bytes() const { return element_count() * sizeof(Native); }
  Native* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

template <typename T>
struct PlatformStream {
  using type = T;
};

template <typename... Ts>
struct Binding {
  template <typename C>
  Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename Fn>
  int To(Fn&&) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "handler signature does not match its binding");
    return 0;
  }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, fn, binding)                 \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame* frame) {             \
    static const int bound = (binding).To(fn);                           \
    (void)bound;                                                         \
    (void)frame;                                                         \
    return nullptr;                                                      \
  }
