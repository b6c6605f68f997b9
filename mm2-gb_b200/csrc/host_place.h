// Host placement next to a GPU.  Every end-to-end path here is bound by host memory passes over pinned staging, so on
// a box with more than one NUMA node the staging pages belong on the node the GPU hangs off: NearGpu confines the CALLING thread to
// the CPUs Linux lists as local to the device (sysfs local_cpulist of its PCI function) for the lifetime of the guard -- pages first
// touched or pinned meanwhile land on that node -- and restores the thread's affinity afterwards.  Nothing happens where the
// kernel shows no topology (the list is missing, empty, or covers every CPU the thread may use), or with MM2GB_NUMA=0.
// The reference allocates its pinned buffers from whichever thread calls init (gpu/plmem.cu:12-143); it has no placement at all.
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>

namespace mm2gb {

// parses "0-31,64-95"; returns the number of CPUs set
inline int parse_cpulist(const char *s, cpu_set_t *set)
{
    CPU_ZERO(set);
    int n = 0;
    const char *p = s;
    while (*p && *p != '\n') {
        char *e = nullptr;
        const long a = strtol(p, &e, 10);
        if (e == p || a < 0) break;
        long b = a;
        if (*e == '-') {
            p = e + 1;
            b = strtol(p, &e, 10);
            if (e == p || b < a) break;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (!CPU_ISSET((int)c, set)) { CPU_SET((int)c, set); ++n; }
        if (*e != ',') break;
        p = e + 1;
    }
    return n;
}

// CPUs local to `device` that the calling thread may use; false when there is nothing to gain from moving there
inline bool gpu_local_cpus(int device, cpu_set_t *out)
{
    const char *env = getenv("MM2GB_NUMA");
    if (env && atoi(env) == 0) return false;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus) - 1, device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char *p = bus; *p; ++p) *p = (char)tolower((unsigned char)*p);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[4096];
    const bool got = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    cpu_set_t local, cur, both;
    if (!got || parse_cpulist(buf, &local) == 0) return false;
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return false;
    CPU_AND(&both, &local, &cur);
    if (CPU_COUNT(&both) == 0 || CPU_EQUAL(&both, &cur)) return false;
    *out = both;
    return true;
}

struct NearGpu {
    cpu_set_t old;
    bool moved = false;
    explicit NearGpu(int device)
    {
        cpu_set_t near;
        if (gpu_local_cpus(device, &near) && sched_getaffinity(0, sizeof(old), &old) == 0)
            moved = sched_setaffinity(0, sizeof(near), &near) == 0;
    }
    ~NearGpu() { if (moved) sched_setaffinity(0, sizeof(old), &old); }
    NearGpu(const NearGpu &) = delete;
    NearGpu &operator=(const NearGpu &) = delete;
};

}  // namespace mm2gb
