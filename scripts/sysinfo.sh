# Host / GPU topology of the GPU box (what the e2e numbers depend on).
nproc; echo "affinity: $(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz|L2|L3' 
free -g | head -2
nvidia-smi topo -m 2>/dev/null | head -20
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
cat /proc/meminfo | egrep 'Huge|MemAvail'
