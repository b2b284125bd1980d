#include "common.cuh"

#include <mutex>

namespace mb {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) fail(B200_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dtype, uint32_t elem_bytes, uint64_t cols,
                         uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows,
                         CUtensorMapSwizzle swizzle) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estride[2] = {1, 1};
    (void)elem_bytes;
    CUresult r = encode_tiled_fn()(&tm, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu pitch=%llu box=%ux%u", (int)r,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_cols,
             box_rows);
    return tm;
}

int sm_count(int device) {
    int n = 0;
    MB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    return n;
}

}  // namespace mb

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

const char* b200_last_error(void) { return mb::g_last_error.c_str(); }

int b200_device_count(int* out_count) {
    return mb::guarded([&] {
        MB_CHECK_ARG(out_count != nullptr, "out_count is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        int ok = 0;
        for (int d = 0; d < n; ++d) {
            int major = 0;
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
                ++ok;
        }
        *out_count = ok;
    });
}

int b200_fuse_vectors(const double* vecs, const double* weights, int n, int dim, int normalize, double* out) {
    return mb::guarded([&] {
        MB_CHECK_ARG(vecs && weights && out, "NULL argument");
        MB_CHECK_ARG(n > 0 && dim > 0, "n and dim must be positive");
        // np.mean([w_i * v_i], axis=0): pairwise order of numpy's add.reduce over axis 0 is sequential
        // for a short leading axis, so a plain left-to-right sum reproduces it.
        for (int d = 0; d < dim; ++d) {
            double acc = 0.0;
            for (int i = 0; i < n; ++i) acc += vecs[(size_t)i * dim + d] * weights[i];
            out[d] = acc / (double)n;
        }
        if (normalize) {
            double ss = 0.0;
            for (int d = 0; d < dim; ++d) ss += out[d] * out[d];
            double nrm = sqrt(ss);
            if (nrm > 0.0)
                for (int d = 0; d < dim; ++d) out[d] /= nrm;
        }
    });
}

}  // extern "C"
