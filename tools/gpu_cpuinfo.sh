#!/bin/bash
echo "nproc $(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -8
grep -c processor /proc/cpuinfo; grep "model name" /proc/cpuinfo | head -1; cat /proc/loadavg
taskset -p $$ 
python - <<'P'
import os
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:40])
P
for spin in 200 2000; do
  export OAT_B200_SPIN_US=$spin
  python - <<'P'
import numpy as np, subprocess, os, time
B='oat_b200/bin/'
np.save('/tmp/img.npy', np.zeros((1080,1920,3),np.uint8))
for rep in range(4):
    subprocess.run([B+'oat-clean','ratetest'],capture_output=True)
    c=subprocess.Popen([B+'shmemdf_test','count-frames','ratetest'],stdout=subprocess.PIPE,text=True)
    time.sleep(0.3)
    subprocess.run([B+'oat-frameserve','test','ratetest','-f','/tmp/img.npy','-n','200000'])
    print('spin', os.environ['OAT_B200_SPIN_US'], 'frames', c.communicate()[0].strip())
P
done
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -8
