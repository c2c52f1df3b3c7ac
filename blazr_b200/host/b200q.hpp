// b200q.hpp -- C++ host mirror of the reference's operator interface above the C ABI (include/b200q.h).
//
// The reference host is Rust (no Rust toolchain in this image), so the compiled-language host side is C++:
// same names and argument meaning as the reference's interface for this path
//   boostr::quant::QuantMatmulOps::quant_matmul / DequantOps::dequantize   (bounds: src/loader/api.rs:25)
//   boostr::quant::decomposed::DecomposedQuantTensor / DecomposedQuantMethod (src/loader/safetensors/awq.rs:218, gptq.rs:252)
//   TensorParallelState::shard_range                                        (src/engine/tensor_parallel.rs:61-67)
// and the same error behaviour: failures surface as a Result-like exception carrying the backend message
// (blazr maps boostr::error / NumrError::Backend(String) to anyhow, src/engine/cuda_graphs.rs:127).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200q.h"

namespace b200q_host {

struct BackendError : std::runtime_error {
    int code;
    BackendError(int c, const std::string& m) : std::runtime_error("b200q error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int32_t rc) {
    if (rc != B200Q_OK) throw BackendError(rc, b200q_last_error());
}

enum class DType : int32_t { F32 = B200Q_F32, F16 = B200Q_F16, BF16 = B200Q_BF16 };

struct DecomposedQuantMethod {
    enum Kind { Awq, Gptq } kind;
    int group_size;
};

// mirror of DecomposedQuantTensor::new(qweight, scales, qzeros, g_idx, method, logical_shape)
struct DecomposedQuantTensor {
    const uint32_t* qweight;   // AWQ [K, N/8] | GPTQ [K/8, N]
    const float* scales;       // [groups, N] (f16-representable)
    const void* qzeros;        // AWQ: f32 [groups, N] (unpacked) | GPTQ: u32 [groups, N/8] (packed)
    const int32_t* g_idx;      // GPTQ only, nullable
    DecomposedQuantMethod method;
    int64_t n, k;              // logical [N, K]
    const float* bias = nullptr;
    int zero_plus_one = 1;
};

// immutable, ref-counted like boostr's Arc storage (src/engine/cuda_graphs.rs:84-85 clones are cheap)
class QuantTensor {
  public:
    QuantTensor() = default;
    explicit QuantTensor(b200q_weight* w) : h_(w, [](b200q_weight* p) { b200q_weight_free(p); }) { check(b200q_weight_info(w, &info_)); }
    const b200q_weight* handle() const { return h_.get(); }
    const b200q_weight_info_t& info() const { return info_; }
    int64_t n() const { return info_.N; }
    int64_t k() const { return info_.K; }

  private:
    std::shared_ptr<b200q_weight> h_;
    b200q_weight_info_t info_{};
};

// reference src/engine/tensor_parallel.rs:22-67
struct TensorParallelState {
    int64_t rank = 0, world_size = 1;
    bool is_root() const { return rank == 0; }
    std::pair<int64_t, int64_t> shard_range(int64_t total, int64_t granule = 1) const {
        int64_t s, e;
        check(b200q_shard_range_blocks(total, granule, rank, world_size, &s, &e));
        return {s, e};
    }
};

// the backend client: QuantMatmulOps + DequantOps on one B200, one stream supplied by the caller
class B200Client {
  public:
    B200Client(int device, void* stream) : device_(device), stream_(stream) {}

    // VarMap::from_gguf upload of one tensor (raw ggml blocks, src/loader/gguf.rs:33)
    QuantTensor weight_from_ggml(int ggml_type, const void* blocks, bool on_device, int64_t n, int64_t k) const {
        b200q_weight* w = nullptr;
        check(b200q_weight_from_ggml(ggml_type, blocks, on_device, n, k, device_, stream_, &w));
        return QuantTensor(w);
    }
    QuantTensor weight_from_ggml_shard(int ggml_type, const void* blocks, bool on_device, int64_t n, int64_t k,
                                       std::pair<int64_t, int64_t> rows, std::pair<int64_t, int64_t> cols) const {
        b200q_weight* w = nullptr;
        check(b200q_weight_from_ggml_shard(ggml_type, blocks, on_device, n, k, rows.first, rows.second, cols.first, cols.second, device_, stream_, &w));
        return QuantTensor(w);
    }
    QuantTensor weight_from_decomposed(const DecomposedQuantTensor& t, bool on_device = false) const {
        b200q_weight* w = nullptr;
        if (t.method.kind == DecomposedQuantMethod::Awq)
            check(b200q_weight_from_awq(t.qweight, t.scales, static_cast<const float*>(t.qzeros), on_device, t.method.group_size, t.n, t.k, device_,
                                        stream_, &w));
        else
            check(b200q_weight_from_gptq(t.qweight, t.scales, static_cast<const uint32_t*>(t.qzeros), t.g_idx, t.bias, on_device, t.method.group_size,
                                         t.zero_plus_one, t.n, t.k, device_, stream_, &w));
        return QuantTensor(w);
    }

    size_t workspace_bytes(const QuantTensor& w, int64_t m) const { return b200q_workspace_bytes(w.handle(), m); }

    // QuantMatmulOps::quant_matmul: y[M,N] = x[M,K] . dequant(w)^T (+bias); graph-capturable
    void quant_matmul(const void* x, DType xdt, int64_t m, int64_t ldx, const QuantTensor& w, void* y, DType ydt, int64_t ldy, void* workspace,
                      size_t workspace_bytes) const {
        check(b200q_matmul(w.handle(), x, static_cast<int32_t>(xdt), m, ldx, y, static_cast<int32_t>(ydt), ldy, workspace, workspace_bytes, stream_));
    }
    // DequantOps::dequantize
    void dequantize(const QuantTensor& w, void* out, DType dt) const { check(b200q_dequantize(w.handle(), out, static_cast<int32_t>(dt), stream_)); }

  private:
    int device_;
    void* stream_;
};

}  // namespace b200q_host
