// exchange.cu -- multi-GPU strip exchange (halo rows, deposit counts, agent migration).
// Placeholder until the NCCL path lands: every entry point fails loudly.
#include "engine.h"

extern "C" int sm_comm_unique_id(uint8_t id[SM_COMM_ID_BYTES])
{
    (void)id;
    return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built into this library yet");
}
extern "C" int sm_comm_init(sm_engine* e, const uint8_t id[SM_COMM_ID_BYTES])
{
    (void)e; (void)id;
    return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built into this library yet");
}
int sm_engine::init_agents_strip(uint64_t) { return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built yet"); }
int sm_engine::exchange_counts() { return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built yet"); }
int sm_engine::exchange_trail_ghosts() { return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built yet"); }
int sm_engine::migrate_agents() { return sm_fail(SM_ERR_STATE, "multi-GPU exchange is not built yet"); }
void sm_engine::comm_destroy() {}
