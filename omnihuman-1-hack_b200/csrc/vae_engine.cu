#include "vae_engine.h"

namespace b2 {
struct VaeEngine::Impl {};
VaeEngine::VaeEngine(int, int) {}
VaeEngine::~VaeEngine() { delete impl; }
void VaeEngine::load_weight(const char*, const void*, int, int, const int64_t*) { fail("VAE engine not built yet"); }
void VaeEngine::finalize() { fail("VAE engine not built yet"); }
void VaeEngine::decode(const float*, int, int, int, float*, cudaStream_t) { fail("VAE engine not built yet"); }
}  // namespace b2
