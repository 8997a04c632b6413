#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own in-tree sources compiled unmodified
(oracle/_ref/librdsp_ref.so, built by oracle/Makefile from /root/reference).  Run in the build container
(the GPU box has no /root/reference); the vectors are committed so that the parity tests can use them
anywhere.

  conv_nr0.npz      FFT-256 overlap-save filter, PBT 300-4000, NR off            (RDSP_convolutional.h:228-353)
  conv_nr30.npz     same + NLMS DNR level 30                                      (RDSP_noise_reduction.h:35-80)
  conv_levels.npz   NR level sequence 0,20,20,50,0,50,30 ... (re-init / stale-ring quirks C6, C7), PBT 200-2800
  conv_kat.npz      the 1 kHz real-tone known-answer test of SURVEY.md section 4.3
  spec256.npz       HP-less AudioAnalyzeFFT256IQ outputs (naverage 30 and 4) on post-biquad style input
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402
from radiodsp_sdr_rx_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    po.build()
    assert po.ref_available(), "oracle/_ref/librdsp_ref.so missing (needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    nb = 48
    iq = synth.synth_iq([101, 202, 303], nb, [0, 1, 4], interferer=[True, False, False])   # [nb,3,128,2]
    # make the levels hot enough that saturation / truncation paths are exercised on one channel
    hot = np.clip(iq[:, 2].astype(np.int32) * 6, -32768, 32767).astype(np.int16)

    r = po.RefChannel()
    L, R = iq[:, 0, :, 0], iq[:, 0, :, 1]
    oL, oR, fL, fR = r.conv(L, R, 0)
    np.savez_compressed(os.path.join(OUT, "conv_nr0.npz"), in_L=L, in_R=R, out_L=oL, out_R=oR, f32_L=fL, f32_R=fR,
                        pbt=np.array([300.0, 4000.0]), mask=r.mask())

    r = po.RefChannel()
    L, R = iq[:, 1, :, 0], iq[:, 1, :, 1]
    oL, oR, fL, fR = r.conv(L, R, 30)
    np.savez_compressed(os.path.join(OUT, "conv_nr30.npz"), in_L=L, in_R=R, out_L=oL, out_R=oR, f32_L=fL, f32_R=fR,
                        pbt=np.array([300.0, 4000.0]), nr_level=np.array([30] * nb))

    r = po.RefChannel()
    r.reinit_filter(200.0, 2800.0)
    levels = ([0] * 3 + [20] * 8 + [50] * 6 + [0] * 4 + [50] * 5 + [30] * 10 + [0] * 2 + [40] * 10)[:nb]
    L, R = hot[:, :, 0], hot[:, :, 1]
    oL, oR, fL, fR = r.conv(L, R, levels)
    np.savez_compressed(os.path.join(OUT, "conv_levels.npz"), in_L=L, in_R=R, out_L=oL, out_R=oR, f32_L=fL, f32_R=fR,
                        pbt=np.array([200.0, 2800.0]), nr_level=np.array(levels), mask=r.mask())

    n = np.arange(41 * 128)
    L = np.rint(16384 * np.sin(2 * np.pi * 1000 * n / 44100)).astype(np.int16).reshape(41, 128)
    R = np.zeros_like(L)
    r = po.RefChannel()
    ci, cq = r.fir()
    oL, oR, fL, fR = r.conv(L, R, 0)
    r2 = po.RefChannel()
    oL2, oR2, _, _ = r2.conv(L, R, 30)
    np.savez_compressed(os.path.join(OUT, "conv_kat.npz"), in_L=L, in_R=R, out_L=oL, out_R=oR, out_L_nr30=oL2, out_R_nr30=oR2,
                        fir_I=ci, fir_Q=cq, mask=r.mask(), mu15=np.float32(r.lms_mu()), mu30=np.float32(r2.lms_mu()))

    nb2 = 70
    iq2 = synth.synth_iq([7, 8], nb2, [0, 4])
    specs = {}
    for nav, ch in ((30, 0), (4, 1)):
        r = po.RefChannel(naverage=nav)
        res = r.fft256(iq2[:, ch, :, 0], iq2[:, ch, :, 1])
        specs[f"idx_nav{nav}"] = np.array([k for k, _ in res])
        specs[f"out_nav{nav}"] = np.stack([o for _, o in res])
        specs[f"in_nav{nav}"] = iq2[:, ch]
    # full-scale square wave: saturating butterflies
    sq = np.where((np.arange(nb2 * 128) // 5) % 2 == 0, 32767, -32768).astype(np.int16).reshape(nb2, 128)
    r = po.RefChannel(naverage=2)
    res = r.fft256(sq, -sq - 1)
    specs["idx_sat"] = np.array([k for k, _ in res])
    specs["out_sat"] = np.stack([o for _, o in res])
    specs["in_sat"] = np.stack([sq, (-sq - 1).astype(np.int16)], axis=-1)
    np.savez_compressed(os.path.join(OUT, "spec256.npz"), **specs)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
