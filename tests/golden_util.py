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


def ci_check(fin_path, ref_path=None, nlayers=2):
    """CI/bump/check.F90 restated: the run is rejected when a layer's mass loss exceeds 1e-12 (check.F90:58-62); for the
    fields of lines 3..5 of each layer block (u, v, ssh) it reports |ref - val| / |ref| of max and min (check.F90:64-80).
    Returns {(layer, field): (err_max, err_min)}."""
    ref_path = ref_path or os.path.join(HERE, "golden", "ci_bump_ref_mlswe_FIN.txt")
    ref = open(ref_path).read().split("\n")
    fin = open(fin_path).read().split("\n")
    out = {}
    for nl in range(1, nlayers + 1):
        index_layer = 6 * (nl - 1) + 1                      # 1-based line numbers, as in the Fortran
        mass_loss = float(fin[index_layer + 1 - 1].split("=")[1])
        assert mass_loss <= 1.0e-12, "Layer %d mass_loss = %g to large" % (nl, mass_loss)
        for ifield in range(2, 5):
            index_field = ifield + index_layer
            r = ref[index_field - 1].split("=")[1].split()
            v = fin[index_field - 1].split("=")[1].split()
            assert r[0] == v[0]
            out[(nl, r[0])] = (abs(float(r[1]) - float(v[1])) / abs(float(r[1])), abs(float(r[2]) - float(v[2])) / abs(float(r[2])))
    return out
