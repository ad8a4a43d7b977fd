// TEST INFRASTRUCTURE ONLY: nothing of csrc/host_utils.h is needed by the kernels themselves (the C-ABI entry points
// that use it are cut off by build_hostsim.py).
#pragma once
