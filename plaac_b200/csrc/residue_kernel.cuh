// Per-residue mode (plotsomefastas, plaac.java:610-647) -- placeholder until the kernels land.
#pragma once
#include "common.cuh"

namespace plaac {

inline int residue_setup(const KScalars&, int) { return PLAAC_OK; }

inline int launch_residue(const KScalars&, const DeviceTables*, const BatchView&, const plaac_residue_out&, int64_t, int,
                          cudaStream_t, int64_t*)
{
    return PLAAC_E_UNSUPPORTED;
}

}  // namespace plaac
