/* Minimal stand-in for <cuda_runtime.h>, only for tests/test_async_harness.py: erlamsa_b200/csrc/eb_async.cpp is host code over the
 * engine's synchronous entry point, so its queueing / threading logic can be compiled and exercised (under ThreadSanitizer) on a
 * box without a GPU, with the engine replaced by tests/mock_cuda/async_harness.cpp. Never on any product include path. */
#ifndef EB200_TEST_MOCK_CUDA_RUNTIME_H
#define EB200_TEST_MOCK_CUDA_RUNTIME_H
typedef int cudaError_t;
typedef struct mock_stream* cudaStream_t;
enum { cudaSuccess = 0, cudaStreamNonBlocking = 1 };
extern "C" {
cudaError_t cudaSetDevice(int device);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamDestroy(cudaStream_t st);
}
#endif
