import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def read_fin(path=None):
    """Parse a mlswe_FIN.txt file (print_diagnostics.F90:167-184) -> {layer: {'mass_loss': x, 'h': (max,min), ...}}"""
    path = path or os.path.join(HERE, "golden", "ci_bump_ref_mlswe_FIN.txt")
    out, cur = {}, None
    for line in open(path):
        m = re.match(r"\s*Layer =\s*(\d+)", line)
        if m:
            cur = int(m.group(1)); out[cur] = {}; continue
        m = re.match(r"\s*Mass Loss\s*=\s*(\S+)", line)
        if m:
            out[cur]["mass_loss"] = float(m.group(1)); continue
        m = re.match(r"\s*Fields:\s*Max/Min = (\w+)\s+(\S+)\s+(\S+)", line)
        if m:
            out[cur][m.group(1)] = (float(m.group(2)), float(m.group(3)))
    return out


def extrema(diag):
    """diag from decks.diagnostics / Oracle.diagnostics -> same structure as read_fin (without mass_loss)."""
    out = {}
    nl = diag["h"].shape[0]
    for k in range(nl):
        out[k + 1] = {f: (float(diag[f][k].max()), float(diag[f][k].min())) for f in ("h", "u", "v", "ssh")}
    return out


def compare_extrema(got, ref):
    """max over entries of |got-ref| scaled by the natural magnitude of the field family (h and ssh share the layer
    thickness scale; u, v share the velocity scale)."""
    worst = 0.0
    for layer, fields in ref.items():
        hscale = max(abs(fields["h"][0]), abs(fields["h"][1]))
        uscale = max(abs(x) for f in ("u", "v") for x in fields[f])
        for f in ("h", "u", "v", "ssh"):
            scale = hscale if f in ("h", "ssh") else uscale
            for a, b in zip(got[layer][f], fields[f]):
                worst = max(worst, abs(a - b) / scale)
    return worst
