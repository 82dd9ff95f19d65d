"""
Generates tests/golden/mpm_golden.json by importing the reference's own mpmath
oracle (reheatfunq/_testing/mpm.py) by path from /root/reference.  Runs only in
the build container (the GPU box has no /root/reference); the JSON is committed.

  python tests/golden/make_mpm_golden.py      (takes a few minutes)
"""
import importlib.util, json, os, sys, tempfile
import numpy as np
from mpmath import mp

HERE = os.path.dirname(os.path.abspath(__file__))
os.chdir(tempfile.mkdtemp())   # mpm.py creates ./.cache-mpm
spec = importlib.util.spec_from_file_location(
    "mpm", "/root/reference/reheatfunq/_testing/mpm.py")
mpm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mpm)

DPS = 30
out = {}

# --- case A: data of testing/test_anomaly_posterior.py:267-279 (default prior) ---
pA = dict(p=2.5220172925228335, s=15.373016724405591, n=0.2184765374862465,
          v=0.2184765374862465, amin=1.0)
qA = np.array([75.70255678, 76.18744121, 79.01339271, 89.41498501, 99.3258167,
               100.00115816, 103.74341636, 106.24710206, 111.40302288, 126.80598511])
cA = np.array([3.74039144e-12, 8.63295937e-12, 3.28571247e-12, 6.02724571e-11,
               3.95429084e-12, 2.38382437e-11, 4.73856217e-12, 1.22700220e-10,
               4.76447511e-11, 1.43570283e-11])
P = mpm.AnomalyPosteriorMPM(qA, cA, pA["p"], pA["s"], pA["n"], pA["v"], pA["amin"], DPS)
mp.dps = DPS
PHmax = float(P.PHmax)
# pdf on a grid that differs from the reference's test nodes, plus cdf by quadrature
xs = [PHmax * f for f in (0.0, 0.013, 0.11, 0.29, 0.47, 0.62, 0.8, 0.9, 0.95, 0.99, 0.999)]
brk = [P.amin, 5, 10, 20, 40, 80, 160, mp.inf]
pdf = [float(mp.quad(lambda a: P.I(mp.mpf(x) / P.PHmax, a), brk) / P.Psi) for x in xs]
cdf = []
for x in xs:
    if x == 0.0:
        cdf.append(0.0)
        continue
    zmax = mp.mpf(x) / P.PHmax
    val = mp.quad(lambda z: mp.quad(lambda a: P.I(z, a), brk), [0, zmax / 2, zmax])
    cdf.append(float(val * P.PHmax / P.Psi))
out["A"] = dict(prior=pA, q=qA.tolist(), c=cA.tolist(), PH=xs, pdf=pdf, cdf=cdf,
                PHmax=PHmax, Psi=float(P.Psi))
print("A done", flush=True)

# --- case B: data of testing/test_anomaly_posterior.py:348-397 (flat prior, z -> 1) ---
qc_i = np.array([
    [74.899297645192660866, 9.0062284112191297855e-09],
    [65.059810425317650129, 6.292979631087474626e-08],
    [73.7333051571991831, 1.7312867537817080333e-07],
    [32.101166178978616017, 3.3383770134386603678e-08],
    [70.933876424473609745, 9.3611019395612739747e-09],
    [27.975606558522464695, 1.0604084288016606936e-08],
    [51.91965658284368601, 5.989265793684179317e-09],
    [66.893346975152468303, 3.3658297443567555587e-09],
    [65.329735427875334608, 1.6897876188526839879e-08],
    [34.859875907287054986, 8.0360194318523393276e-09]])
qi = qc_i[:, 0].copy()
ci = qc_i[:, 1].copy()
qB = qi + 100e6 * ci
P = mpm.AnomalyPosteriorMPM(qB, ci, 1.0, 0.0, 0.0, 0.0, 1.0, DPS)
mp.dps = DPS
PHmax = float(P.PHmax)
# distances from PHmax (relative y) spanning the regular and the large-z branch
ys = [0.5, 0.1, 1e-2, 1e-3, 1e-4, 1e-5, 6e-6, 1e-6, 1e-7, 1e-8, 1e-10, 1e-12]
brk = [P.amin, 5, 10, 20, 40, 80, mp.inf]
pdfB = []
for y in ys:
    z = mp.mpf(1) - mp.mpf(y)
    pdfB.append(float(mp.quad(lambda a: P.I(z, a), brk) / P.Psi))
out["B"] = dict(prior=dict(p=1.0, s=0.0, n=0.0, v=0.0, amin=1.0), q=qB.tolist(),
                c=ci.tolist(), y=ys, pdf=pdfB, PHmax=PHmax, Psi=float(P.Psi))
print("B done", flush=True)

with open(os.path.join(HERE, "mpm_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
