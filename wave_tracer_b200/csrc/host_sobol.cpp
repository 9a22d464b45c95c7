// host_sobol.cpp -- wthost_sobol_tables (include/wthost.h): the sobolld generator matrices as the device keeps them, for tests and tools.
// Part of libwt_host.so (host-only code: no CUDA runtime, nothing of the device path).
#include "sobol_tables.h"
#include "../../include/wthost.h"
extern "C" int wthost_sobol_tables(const wtgpu_sobol_entry* table, uint16_t* ones, uint16_t* twos) {
    if (!table || !ones || !twos) return WTGPU_E_INVALID;
    std::string why;
    return wt::sobol_build_tables(table, reinterpret_cast<uint16_t (*)[WTGPU_SOBOL_DIGITS]>(ones), reinterpret_cast<uint16_t (*)[WTGPU_SOBOL_DIGITS]>(twos), why) ? WTGPU_OK : WTGPU_E_INVALID;
}
