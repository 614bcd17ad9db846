// Narrow window onto b200w_ctx (defined in engine.cu) for the other translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct b200w_ctx;

int ctx_device(b200w_ctx* c);
cudaStream_t ctx_stream(b200w_ctx* c);
void ctx_set_error(b200w_ctx* c, const char* msg);
int64_t& ctx_launches(b200w_ctx* c);
void* ctx_infer_slot(b200w_ctx* c);
void ctx_set_infer(b200w_ctx* c, void* p, void (*destroy)(void*));
// normal(0, std) / constant fill of a bf16 range with the engine's counter-based generator
void ctx_fill_normal(b200w_ctx* c, void* w_bf16, size_t n, uint64_t seed, float std);
void ctx_fill_const(b200w_ctx* c, void* w_bf16, size_t n, float value);
