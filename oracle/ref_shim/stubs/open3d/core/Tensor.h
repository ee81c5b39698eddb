// ref_shim stub: the few members of core::Tensor / Dtype / Device / SizeVector the
// reference's *Impl.h headers touch, over caller-owned host memory.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <initializer_list>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "open3d/utility/Logging.h"

namespace open3d {
namespace core {

class SizeVector : public std::vector<int64_t> {
public:
    using std::vector<int64_t>::vector;
};

class Dtype {
public:
    Dtype() : size_(0), code_(0) {}
    Dtype(int64_t size, int code) : size_(size), code_(code) {}
    int64_t ByteSize() const { return size_; }
    bool operator==(const Dtype& o) const { return size_ == o.size_ && code_ == o.code_; }
    bool operator!=(const Dtype& o) const { return !(*this == o); }
    std::string ToString() const { return "Dtype"; }
    static const Dtype Float32, Float64, UInt8, UInt16, Int32, Int64, Bool, Int8, Int16, UInt32, UInt64;
private:
    int64_t size_;
    int code_;
};
inline const Dtype Dtype::Float32(4, 1);
inline const Dtype Dtype::Float64(8, 2);
inline const Dtype Dtype::UInt8(1, 3);
inline const Dtype Dtype::UInt16(2, 4);
inline const Dtype Dtype::Int32(4, 5);
inline const Dtype Dtype::Int64(8, 6);
inline const Dtype Dtype::Bool(1, 7);
inline const Dtype Dtype::Int8(1, 8);
inline const Dtype Dtype::Int16(2, 9);
inline const Dtype Dtype::UInt32(4, 10);
inline const Dtype Dtype::UInt64(8, 11);
static const Dtype Float32 = Dtype::Float32;
static const Dtype Float64 = Dtype::Float64;
static const Dtype UInt8 = Dtype::UInt8;
static const Dtype UInt16 = Dtype::UInt16;
static const Dtype Int32 = Dtype::Int32;
static const Dtype Int64 = Dtype::Int64;
static const Dtype Bool = Dtype::Bool;
static const Dtype Int8 = Dtype::Int8;
static const Dtype Int16 = Dtype::Int16;
static const Dtype UInt32 = Dtype::UInt32;
static const Dtype UInt64 = Dtype::UInt64;

class Device {
public:
    Device() {}
    explicit Device(const std::string& s) : cuda_(s.rfind("CUDA", 0) == 0) {}
    bool operator==(const Device& o) const { return cuda_ == o.cuda_; }
    bool IsCPU() const { return !cuda_; }
    bool IsCUDA() const { return cuda_; }
    bool IsSYCL() const { return false; }
    std::string ToString() const { return cuda_ ? "CUDA:0" : "CPU:0"; }
private:
    bool cuda_ = false;   // only integration/ (built with O3DB_STUB_TENSOR_CUDA) ever creates CUDA stub tensors
};

class Tensor {
public:
    Tensor() : ptr_(nullptr) {}
    // non-owning view of caller memory (host, or device memory when `device` says so)
    Tensor(void* ptr, SizeVector shape, Dtype dtype, const Device& device = Device())
        : ptr_(ptr), shape_(shape), dtype_(dtype), device_(device) {}
    // owning, zero-initialised (upstream: uninitialised)
    Tensor(const SizeVector& shape, Dtype dtype, const Device& device = Device()) : shape_(shape), dtype_(dtype), device_(device) {
        int64_t n = 1;
        for (auto s : shape_) n *= s;
        const size_t bytes = (size_t)(n > 0 ? n : 1) * dtype.ByteSize();
#ifdef O3DB_STUB_TENSOR_CUDA
        if (device.IsCUDA()) {   // integration/forwarder_hooks.cpp: outputs the forwarders allocate live on the device
            void* d = nullptr;
            if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemset(d, 0, bytes) != cudaSuccess)
                utility::LogError("ref_shim Tensor: cudaMalloc failed");
            own_ = std::shared_ptr<char>(static_cast<char*>(d), [](char* q) { cudaFree(q); });
            ptr_ = d;
            return;
        }
#endif
        own_ = std::shared_ptr<char>(new char[bytes](), std::default_delete<char[]>());
        ptr_ = own_.get();
    }
    template <typename T>
    Tensor(const std::vector<T>& init, const SizeVector& shape, Dtype dtype, const Device& d = Device())
        : Tensor(shape, dtype, d) {
        memcpy(ptr_, init.data(), init.size() * sizeof(T));
    }
    static Tensor Zeros(const SizeVector& shape, Dtype dtype, const Device& d = Device()) { return Tensor(shape, dtype, d); }
    static Tensor Empty(const SizeVector& shape, Dtype dtype, const Device& d = Device()) { return Tensor(shape, dtype, d); }
    static Tensor Eye(int64_t n, Dtype, const Device&) {
        static double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        return Tensor(eye, {n, n}, Dtype::Float64);
    }
    bool IsContiguous() const { return true; }
    SizeVector GetShape() const { return shape_; }
    int64_t GetShape(int i) const { return shape_[i]; }
    int64_t NumDims() const { return (int64_t)shape_.size(); }
    int64_t GetLength() const { return shape_.empty() ? 0 : shape_[0]; }
    int64_t NumElements() const {
        if (ptr_ == nullptr) return 0;
        int64_t n = 1;
        for (auto s : shape_) n *= s;
        return n;
    }
    Dtype GetDtype() const { return dtype_; }
    Device GetDevice() const { return device_; }
    void* GetDataPtr() { return ptr_; }
    const void* GetDataPtr() const { return ptr_; }
    template <typename T>
    T* GetDataPtr() { return static_cast<T*>(ptr_); }
    template <typename T>
    const T* GetDataPtr() const { return static_cast<const T*>(ptr_); }
    template <typename T>
    T Item() const { return *static_cast<const T*>(ptr_); }
    // sub-tensor along dim 0 (views share ownership)
    Tensor operator[](int64_t i) const {
        SizeVector sub(shape_.begin() + 1, shape_.end());
        int64_t stride = dtype_.ByteSize();
        for (auto s : sub) stride *= s;
        Tensor t(static_cast<char*>(ptr_) + i * stride, sub, dtype_, device_);
        t.own_ = own_;
        return t;
    }
    Tensor Slice(int64_t dim, int64_t start, int64_t stop) const {
        if (dim != 0) utility::LogError("ref_shim Tensor::Slice: dim 0 only");
        SizeVector shp = shape_;
        shp[0] = stop - start;
        int64_t stride = dtype_.ByteSize();
        for (size_t k = 1; k < shape_.size(); ++k) stride *= shape_[k];
        Tensor t(static_cast<char*>(ptr_) + start * stride, shp, dtype_, device_);
        t.own_ = own_;
        return t;
    }
    // ---- members that only have to COMPILE (ComputeRtPointToPointCPU etc. are never run in the shim)
    Tensor To(const Dtype&) const { unsupported(); }
    Tensor To(const Device&, const Dtype& dtype = Dtype()) const {
        if (dtype == Dtype::Float64 && dtype_ == Dtype::Float32) {   // the one conversion the compiled kernels ask for
            Tensor t(shape_, Dtype::Float64);
            const float* in = static_cast<const float*>(ptr_);
            double* out = static_cast<double*>(t.ptr_);
            for (int64_t i = 0; i < NumElements(); ++i) out[i] = in[i];
            return t;
        }
        return *this;
    }
    std::tuple<Tensor, Tensor, Tensor> SVD() const { unsupported(); }
    double Det() const { unsupported(); }
    Tensor T() const { unsupported(); }
    Tensor Matmul(const Tensor&) const { unsupported(); }
    Tensor Reshape(const SizeVector&) const { unsupported(); }
    Tensor Neg() const { unsupported(); }
    Tensor Div(double) const { unsupported(); }
    Tensor Transpose(int64_t, int64_t) const { unsupported(); }
    template <typename T>
    static Tensor Full(const SizeVector&, T, Dtype, const Device& = Device()) { unsupported(); }
    Tensor Contiguous() const { return *this; }
    Tensor Clone() const {
        Tensor t(shape_, dtype_);
        memcpy(t.ptr_, ptr_, (size_t)NumElements() * dtype_.ByteSize());
        return t;
    }
    Tensor operator-(const Tensor&) const { unsupported(); }
    Tensor& operator=(double) { unsupported(); }
    Tensor(const Tensor&) = default;
    Tensor& operator=(const Tensor&) = default;
private:
    [[noreturn]] static void unsupported() { utility::LogError("ref_shim Tensor: member not available in the stub"); }
    void* ptr_;
    SizeVector shape_;
    Dtype dtype_;
    Device device_;
    std::shared_ptr<char> own_;
};

}  // namespace core
}  // namespace open3d
