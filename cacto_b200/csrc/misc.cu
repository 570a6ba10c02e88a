// misc.cu -- ABI version and parameter-block sizes.
#include "common.cuh"
#include "mlp_layout.cuh"

extern "C" int32_t cacto_abi_version(void) { return CACTO_ABI_VERSION; }
extern "C" int64_t cacto_actor_param_count(int32_t ns, int32_t na) { return cacto::ActorLayout(ns, na).total; }
extern "C" int64_t cacto_critic_param_count(int32_t ns) { return cacto::CriticLayout(ns).total; }
