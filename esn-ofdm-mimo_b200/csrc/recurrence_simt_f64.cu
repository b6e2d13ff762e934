// fp64 instantiations of the SIMT reservoir recurrence (see recurrence_simt.cuh)
#include "recurrence_simt.cuh"
int esn_simt_launch_f64(const esn_simt::RecParams &p, cudaStream_t st) { return esn_simt::launch_any<double>(p, st); }
