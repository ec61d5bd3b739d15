#!/bin/bash
cd $GRAFT_REPO_ROOT
python - <<'P' > gpurun_out/numa.log 2>&1
import torch, os, glob
pr = torch.cuda.get_device_properties(0)
print([a for a in dir(pr) if 'pci' in a], getattr(pr,'pci_bus_id',None), getattr(pr,'pci_device_id',None), getattr(pr,'pci_domain_id',None))
for f in glob.glob('/sys/bus/pci/devices/*/numa_node')[:400]:
    v=open(f).read().strip()
    if v!='-1' and v!='0': print(f,v)
print('nodes', glob.glob('/sys/devices/system/node/node*'))
print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '...')
P
nvidia-smi topo -m >> gpurun_out/numa.log 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/numa.log 2>&1
cat gpurun_out/numa.log
