// Host placement for one-process-per-GPU runs: keep the calling thread (and the threads / pinned buffers it
// creates afterwards) on the NUMA node the GPU's PCIe root port hangs off, so that H2D / D2H traffic of the
// end-to-end path does not cross the socket interconnect.  No libnuma in the image: sysfs + raw syscalls.
// Nothing here is required for correctness; every failure is reported as "not bound" and the run goes on.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <string>

#include "../../include/cutseq_b200.h"

void csq_set_error(const char* msg);

namespace {

cpu_set_t g_saved_affinity;
bool g_have_saved = false;

bool read_line(const std::string& path, char* buf, size_t cap) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    const bool ok = fgets(buf, (int)cap, f) != nullptr;
    fclose(f);
    return ok;
}

// "0-31,64-95" -> cpu set; false when empty / unparsable
bool parse_cpulist(const char* s, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    while (*s && *s != '\n') {
        char* end;
        long lo = strtol(s, &end, 10);
        if (end == s) return false;
        long hi = lo;
        s = end;
        if (*s == '-') {
            hi = strtol(s + 1, &end, 10);
            if (end == s + 1) return false;
            s = end;
        }
        for (long c = lo; c <= hi && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, set);
            n++;
        }
        if (*s == ',') s++;
    }
    return n > 0;
}

}  // namespace

extern "C" {

int csq_bind_host_to_device(int device, int* numa_node) {
    if (numa_node) *numa_node = -1;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        csq_set_error("csq_bind_host_to_device: no PCI bus id for this device");
        return CSQ_ERR_NO_DEVICE;
    }
    for (char* p = bus; *p; p++)
        if (*p >= 'A' && *p <= 'F') *p = (char)(*p + 32);  // sysfs spells the address in lower case
    char buf[4096];
    if (!read_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", buf, sizeof(buf))) return 0;
    const int node = atoi(buf);
    if (node < 0) return 0;  // single-node machine or the firmware does not say
    if (!read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", buf, sizeof(buf))) return 0;
    cpu_set_t want, have, both;
    if (!parse_cpulist(buf, &want)) return 0;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return 0;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return 0;  // the container's cpuset has no CPU of that node
    if (!g_have_saved) {
        g_saved_affinity = have;
        g_have_saved = true;
    }
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return 0;
    // pages this thread touches (and pins) from now on: prefer that node; MPOL_PREFERRED = 1
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof(mask) * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        (void)syscall(SYS_set_mempolicy, 1, mask, sizeof(mask) * 8);  // EPERM under seccomp is fine: first touch is local anyway
    }
    if (numa_node) *numa_node = node;
    return 0;
}

int csq_unbind_host(void) {
    if (g_have_saved) sched_setaffinity(0, sizeof(g_saved_affinity), &g_saved_affinity);
    (void)syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
    return 0;
}

}  // extern "C"
