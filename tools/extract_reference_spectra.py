#!/usr/bin/env python
"""Extracts the tabulated physical data the benchmark scenes name from the reference's data/ directory into a small fixture
(wave_tracer_b200/data/spectra.npz) so that the scenes can be built where /root/reference is not mounted (the GPU box).

What is read (public physical data, not code): refractive indices of data/ior/*.yml (refractiveindex.info, CC0: tabulated n,k or the
Sellmeier coefficients of "formula 2", evaluated as src/spectrum/util/spectrum_from_db.cpp:86-110 does), the emission spectrum
data/emission/2534_CFL_Tensor_Twister.yml (LSPDD) that scenes/cornell-box/box.xml:277 names, and the CIE XYZ colour-matching functions of
data/sensitivity/XYZ.yml behind <response type="RGB"> (src/sensor/response/RGB.cpp).  Run once in the container that mounts the reference:
    python tools/extract_reference_spectra.py
"""
import os
import sys
import numpy as np
import yaml

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wave_tracer_b200", "data", "spectra.npz")
MATERIALS = ["Au", "Al", "Ag", "Cu", "SF5", "SF11", "BK7"]
EMITTERS = ["2534_CFL_Tensor_Twister", "2723_LED_Greatwall-Ledlight_A19"]


def table(text, ncol):
    rows = [[float(x) for x in ln.split()] for ln in text.strip().splitlines() if ln.strip()]
    return np.array([r[:ncol] for r in rows if len(r) >= ncol], np.float64)


def load_ior(name):
    """-> wavelengths (um) ascending, n, k on the union of the entries' grids"""
    db = yaml.safe_load(open(os.path.join(REF, "ior", name + ".yml")))
    n_tab = k_tab = None
    for e in db["DATA"]:
        t = e["type"].strip()
        if t == "tabulated nk":
            a = table(e["data"], 3); n_tab, k_tab = a[:, [0, 1]], a[:, [0, 2]]
        elif t == "tabulated n":
            n_tab = table(e["data"], 2)
        elif t == "tabulated k":
            k_tab = table(e["data"], 2)
        elif t in ("formula 1", "formula 2"):
            c = [float(x) for x in str(e["coefficients"]).split()]; c += [0.0] * (7 - len(c))
            l1, l2 = [float(x) for x in str(e["wavelength_range"]).split()]
            A, B1, C1, B2, C2, B3, C3 = c[:7]
            if t == "formula 1": C1, C2, C3 = C1 * C1, C2 * C2, C3 * C3
            lam = np.linspace(l1, l2, 256); l2_ = lam * lam
            n2 = 1 + A + B1 * l2_ / (l2_ - C1) + B2 * l2_ / (l2_ - C2) + B3 * l2_ / (l2_ - C3)
            n_tab = np.stack([lam, np.sqrt(np.maximum(n2, 0))], 1)
    lam = n_tab[:, 0]
    k = np.interp(lam, k_tab[:, 0], k_tab[:, 1], left=0, right=0) if k_tab is not None else np.zeros_like(lam)
    o = np.argsort(lam)
    return lam[o], n_tab[o, 1], k[o]


def main():
    out = {}
    for m in MATERIALS:
        lam, n, k = load_ior(m)
        out[f"ior/{m}/lam_um"], out[f"ior/{m}/n"], out[f"ior/{m}/k"] = lam.astype(np.float32), n.astype(np.float32), k.astype(np.float32)
        print(m, len(lam), "points", lam[0], "..", lam[-1], "um; n(0.55)=", np.interp(.55, lam, n), "k(0.55)=", np.interp(.55, lam, k))
    for e in EMITTERS:
        db = yaml.safe_load(open(os.path.join(REF, "emission", e + ".yml")))
        a = table(db["DATA"][0]["data"], 2)
        out[f"emission/{e}/lam_nm"], out[f"emission/{e}/value"] = a[:, 0].astype(np.float32), a[:, 1].astype(np.float32)
        print(e, len(a), "points", a[0, 0], "..", a[-1, 0], "nm")
    db = yaml.safe_load(open(os.path.join(REF, "sensitivity", "XYZ.yml")))
    a = table(db["DATA"][0]["data"], 4)
    out["XYZ/lam_nm"], out["XYZ/xyz"] = a[:, 0].astype(np.float32), a[:, 1:4].astype(np.float32)
    print("XYZ", len(a), "points", a[0, 0], "..", a[-1, 0], "nm")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference data not mounted")
    main()
