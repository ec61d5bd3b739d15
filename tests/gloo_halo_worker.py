"""Worker for test_gloo_halo_ordering_world2 (CPU, gloo): exercises the partition description exactly the way the
NCCL path of the library consumes it (one message per neighbour, faces in nbh_send_recv order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnumo_loader import hnumo_b200 as hn  # noqa: E402


def face_nodes(ilocal, ngl):
    n = np.arange(ngl)
    return {3: n, 4: (ngl - 1) * ngl + n, 5: n * ngl, 6: n * ngl + ngl - 1}[ilocal]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    p = dict(hn.decks.SHIPPED["double_gyre"], nelx=int(os.environ.get("HALO_NELX", "5")), nely=int(os.environ.get("HALO_NELY", "6")),
             partition=os.environ.get("HALO_PARTITION", "rows"))
    d = hn.decks.build_deck(p, rank, world)
    ngl, npts = d["ngl"], d["npts"]
    send, off = [], 0
    reqs, recvs = [], []
    for nb, cnt in zip(d["nbh_proc"], d["num_send_recv"]):
        buf = []
        for i in range(cnt):
            f = d["face"][d["nbh_send_recv"][off + i] - 1]
            nodes = (f[6] - 1) * npts + face_nodes(int(f[4]), ngl)
            buf.append(d["coord"][nodes])  # (ngl, 2)
        off += cnt
        t = torch.from_numpy(np.ascontiguousarray(np.stack(buf)))
        r = torch.empty_like(t)
        reqs.append(dist.isend(t, int(nb) - 1)); reqs.append(dist.irecv(r, int(nb) - 1))
        send.append(t); recvs.append(r)
    for q in reqs:
        q.wait()
    for t, r in zip(send, recvs):
        assert torch.allclose(t, r, rtol=0, atol=1e-9), "halo faces do not pair up"
    print("HALO_OK rank", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
