#!/bin/bash
# nvcc -Xptxas -v of the library: registers / spills per kernel (run from the repo root)
cd "$(dirname "$0")/.." && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -diag-suppress 177,550 -Xptxas -v $EXTRA -o ${OUT:-h-numo_b200/libhnumo_b200.so} h-numo_b200/csrc/hnumo_b200.cu -lpthread -ldl 2>&1 | \
  python3 -c "
import sys,re
name=None
for ln in sys.stdin:
    if 'error' in ln or 'rror:' in ln: print(ln.rstrip())
    m=re.search(r\"Compiling entry function '(\S+)'\",ln)
    if m: name=m.group(1)
    m=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores',ln)
    if m: st=(m.group(1),m.group(2))
    m=re.search(r'Used (\d+) registers',ln)
    if m and name: print('%-90s regs=%s stack=%s spill=%s'%(name[:90],m.group(1),st[0],st[1])); name=None
"
