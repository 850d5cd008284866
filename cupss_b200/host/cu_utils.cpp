// cu_utils.cpp -- fatal CUDA error helper and device listing (semantics of /root/reference/src/cu_utils.cpp).
#include <cstdlib>
#include "../../inc/cupss.h"

void check_error(cudaError_t err) {
    if (err == cudaSuccess) return;
    std::cerr << "CUDA Runtime Error" << std::endl << cudaGetErrorString(err) << std::endl;
    std::exit(1);
}

void check_device() {
    int n = 0, rt = 0, drv = 0;
    check_error(cudaGetDeviceCount(&n));
    check_error(cudaRuntimeGetVersion(&rt));
    check_error(cudaDriverGetVersion(&drv));
    if (n == 0) {
        std::cerr << "No CUDA devices found but trying to run on CUDA, exiting." << std::endl;
        std::exit(1);
    }
    std::cout << "Devices found:" << std::endl;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, i);
        std::cout << "Device " << i << ": " << prop.name << std::endl;
    }
    std::cout << "CUDA Driver Version: " << drv / 1000 << "." << (drv % 100) / 10 << std::endl;
    std::cout << "CUDA Runtime Version: " << rt / 1000 << "." << (rt % 100) / 10 << std::endl;
    if (rt > drv) std::cout << "WARNING: runtime version is not supported by driver. Solver might not work properly" << std::endl;
}
