"""tests/parity_report.py -- parity statistics of the CUDA path against the CPU oracle on 2^20-instance samples of
every BASELINE configuration (run on the B200 box; writes gpurun_out/parity_report.json, copied to profiles/).

For each quantity: max and 99.9th percentile of the per-instance inf-norm relative error |gpu - ref|_inf / |ref|_inf,
and the share of instances inside the flat 1e-12 bar (north_star).  Solves are reported both raw and divided by the
condition number (SURVEY.md sections 0.5 / 7).  The oracle is the C restatement pinned bit-for-bit to the reference.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from matrix_b200 import batch as mb, synth  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

N = 1 << 20


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stats(err):
    err = np.asarray(err)
    return {"max": float(err.max()), "p99.9": float(np.percentile(err, 99.9)), "median": float(np.median(err)),
            "share_below_1e-12": float((err < 1e-12).mean())}


def rel(a, b):
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    den = np.abs(b).max(axis=1)
    return np.abs(a - b).max(axis=1) / np.where(den == 0, 1.0, den)


def main():
    o = Oracle("port")
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    rep = {"instances_per_config": N, "oracle": "oracle/ppx_oracle.c (bit-identical to the reference headers)"}

    xi = synth.se3_twists(N)
    T = o.se3_exp(xi)
    rep["cfg1.se3_exp"] = stats(rel(mb.se3_exp(dev(xi)).cpu().numpy(), T))
    rep["cfg1.SE3_log (same input)"] = stats(rel(mb.SE3_log(dev(T)).cpu().numpy(), o.SE3_log(T)))
    rep["cfg1.exp_log fused round trip"] = stats(rel(mb.se3_exp_log(dev(xi)).cpu().numpy(), o.se3_exp_log(xi)))

    q = synth.joint_angles(N)
    kin = mb.Kinematics(S, M)
    Tg, Jg = kin.fk_jacobian(dev(q))
    Tr, Jr = o.poe_fk_jacobian(S, M, q)
    rep["cfg2.forwardSpace"] = stats(rel(Tg.cpu().numpy(), Tr))
    rep["cfg2.jacobiSpace"] = stats(rel(Jg.cpu().numpy(), Jr))

    A, b = synth.lstsq_4x3(N)
    x, st = mb.linsolve(mb.QR, 4, 3, dev(A), dev(b))
    rx, rst = o.linsolve_qr(4, 3, A, b)
    kappa = np.linalg.cond(A.reshape(-1, 3, 4).transpose(0, 2, 1))
    e = rel(x.cpu().numpy(), rx)
    rep["cfg3.linsolve_qr x"] = stats(e)
    rep["cfg3.linsolve_qr x / kappa"] = stats(e / kappa)
    rep["cfg3.status mismatches"] = int((st.cpu().numpy() != rst).sum())

    A6 = synth.dense(N, 6, 6)
    U, w, V = mb.svdcmp(6, 6, dev(A6))
    _, rw, _ = o.svd(6, 6, A6)
    w = w.cpu().numpy()
    rep["cfg4a.singular values (sorted, rel. to largest)"] = stats(
        np.abs(np.sort(w, axis=1) - np.sort(rw, axis=1)).max(axis=1) / rw.max(axis=1))
    Um = U.cpu().numpy().reshape(N, 6, 6).transpose(0, 2, 1)
    Vm = V.cpu().numpy().reshape(N, 6, 6).transpose(0, 2, 1)
    Am = A6.reshape(N, 6, 6).transpose(0, 2, 1)
    rep["cfg4a.reconstruction |U w V^T - A| / |A|"] = stats(
        np.abs(np.einsum("nij,nj,nkj->nik", Um, w, Vm) - Am).max(axis=(1, 2)) / np.abs(Am).max(axis=(1, 2)))

    S4 = synth.symmetric(N, 4)
    d, z = mb.eig(4, dev(S4))
    rd, _ = o.eig_sym(4, S4, True)
    d = d.cpu().numpy()
    rep["cfg4b.eigenvalues (rel. to largest)"] = stats(np.abs(d - rd).max(axis=1) / np.abs(rd).max(axis=1))
    Zm = z.cpu().numpy().reshape(N, 4, 4).transpose(0, 2, 1)
    Sm = S4.reshape(N, 4, 4)
    rep["cfg4b.residual |S z - z d| / |S|"] = stats(
        np.abs(np.einsum("nij,njk->nik", Sm, Zm) - Zm * d[:, None, :]).max(axis=(1, 2)) / np.abs(Sm).max(axis=(1, 2)))

    qi = synth.joint_angles(N, seed=synth.SEEDS["ik_newton_step"])
    Tt = o.poe_fk_jacobian(S, M, qi + synth.ik_offsets(N), want_Js=False)[0]
    qo, Vs = kin.ik_newton_step(dev(qi), dev(Tt), want_Vs=True)
    rq, rV = o.ik_newton_step(S, M, qi, Tt)
    Js = o.poe_fk_jacobian(S, M, qi, want_T=False)[1]
    sv = np.linalg.svd(Js.reshape(N, 6, 6), compute_uv=False)
    kap = sv[:, 0] / sv[:, -1]
    dq, rdq = qo.cpu().numpy() - qi, rq - qi
    e = np.abs(dq - rdq).max(axis=1) / np.abs(rdq).max(axis=1)
    rep["cfg5.Vs"] = stats(rel(Vs.cpu().numpy(), rV))
    rep["cfg5.dq"] = stats(e)
    rep["cfg5.dq / kappa(Js)"] = stats(e / kap)
    rep["cfg5.dq where kappa(Js) < 1e3"] = stats(e[kap < 1e3])

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(rep, f, indent=1)
    for k, v in rep.items():
        print(k, v)


if __name__ == "__main__":
    main()
