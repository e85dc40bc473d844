// Error plumbing, device selection and checkpoint-tensor upload for libfsb.
#include "fsb_common.cuh"

namespace fsb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libfsb has no CPU fallback",
                  e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return FSB_ERR_CUDA;
    }
    FSB_REQUIRE(device >= 0 && device < n, FSB_ERR_INVALID, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    FSB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    FSB_REQUIRE(prop.major == 10, FSB_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; libfsb is built for sm_100a only and has no fallback path", device,
                prop.major, prop.minor);
    FSB_CUDA_OK(cudaSetDevice(device));
    return FSB_OK;
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16 *__restrict__ in, float *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

static const fsb_tensor *find_tensor(const fsb_tensor *table, size_t n, const std::string &name) {
    for (size_t i = 0; i < n; ++i)
        if (table[i].name && name == table[i].name) return &table[i];
    return nullptr;
}

int upload_tensor(const fsb_tensor *table, size_t n, const std::string &name, std::vector<int64_t> shape,
                  int want_dtype, cudaStream_t stream, DevTensor *out, std::vector<void *> *owned) {
    const fsb_tensor *t = find_tensor(table, n, name);
    FSB_REQUIRE(t != nullptr, FSB_ERR_MISSING_WEIGHT, "checkpoint has no tensor named '%s'", name.c_str());
    FSB_REQUIRE(t->data != nullptr, FSB_ERR_INVALID, "tensor '%s' has a null data pointer", name.c_str());
    FSB_REQUIRE(t->dtype == FSB_F32 || t->dtype == FSB_BF16, FSB_ERR_SHAPE, "tensor '%s': dtype %d not f32/bf16",
                name.c_str(), t->dtype);
    FSB_REQUIRE(t->ndim == (int)shape.size(), FSB_ERR_SHAPE, "tensor '%s': rank %d, expected %zu", name.c_str(),
                t->ndim, shape.size());
    size_t numel = 1;
    for (size_t i = 0; i < shape.size(); ++i) {
        FSB_REQUIRE(t->shape[i] == shape[i], FSB_ERR_SHAPE, "tensor '%s': dim %zu is %lld, expected %lld",
                    name.c_str(), i, (long long)t->shape[i], (long long)shape[i]);
        numel *= (size_t)shape[i];
    }
    const size_t src_es = t->dtype == FSB_F32 ? 4 : 2, dst_es = want_dtype == FSB_F32 ? 4 : 2;
    void *dst = nullptr;
    FSB_CUDA_OK(cudaMalloc(&dst, numel * dst_es));
    owned->push_back(dst);
    const cudaMemcpyKind kind = t->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (t->dtype == want_dtype) {
        FSB_CUDA_OK(cudaMemcpyAsync(dst, t->data, numel * src_es, kind, stream));
        FSB_CUDA_OK(cudaStreamSynchronize(stream));
    } else {
        void *tmp = nullptr;
        FSB_CUDA_OK(cudaMalloc(&tmp, numel * src_es));
        cudaError_t e = cudaMemcpyAsync(tmp, t->data, numel * src_es, kind, stream);
        if (e == cudaSuccess) {
            const int blocks = (int)std::min<size_t>((numel + 255) / 256, 148 * 16);
            if (want_dtype == FSB_BF16)
                f32_to_bf16_kernel<<<blocks, 256, 0, stream>>>((const float *)tmp, (__nv_bfloat16 *)dst, numel);
            else
                bf16_to_f32_kernel<<<blocks, 256, 0, stream>>>((const __nv_bfloat16 *)tmp, (float *)dst, numel);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(tmp);
        FSB_CUDA_OK(e);
    }
    out->ptr = dst;
    out->dtype = want_dtype;
    out->ndim = (int)shape.size();
    for (size_t i = 0; i < shape.size(); ++i) out->shape[i] = shape[i];
    return FSB_OK;
}

}  // namespace fsb

extern "C" {
int fsb_abi_version(void) { return FSB_ABI_VERSION; }
const char *fsb_last_error(void) { return fsb::get_error(); }
int fsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}
}
