#!/bin/bash
# host topology of a multi-GPU box: which NUMA node every GPU hangs off, which cores / memory nodes this container may use
nvidia-smi topo -m 2>&1 | head -40
echo "== numa nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done
echo "== gpus"; for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$d numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) local_cpus $(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"; done
echo "== this container"; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
python - <<'P'
import ctypes
try:
    l=ctypes.CDLL("libnuma.so.1"); print("libnuma: yes", l.numa_available())
except OSError as e: print("libnuma: no")
P
