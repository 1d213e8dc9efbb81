"""Exhaustive check (all 2^32 float bit patterns) of the direction gates of the second-generation scan loop
(scan_pixel_lane2 in csrc/sdm_kernels.cuh) as functions of the RAW angle difference d = gth - c, i.e. BEFORE the
reference's wrap steps (ProbabilityMapping.cc:795-809):

  reference, per gate:   a = d;  if (a >= 360) a -= 360;  if (a < 0) a += 360;   then the fold sequence
      gate 2:  if (a > 180) a = 360 - a;  if (a > 90) a = 180 - a;  reject iff a >= 80
      gate 3:  if (a > 180) a = 360 - a;                            reject iff a >= 45

  scan_pixel_lane2:
      gate 2:  a = d < 0 ? d + 360 : d;                  reject iff  | |a - 180| - 90 | <= 10
      gate 3:  reject iff  (d >= 45 || d <= T2) && |d| <= 315,     T2 = -(45 - 2^-16) = -0x1.67fff8p+5

The new forms drop the `a >= 360` step, so they are claimed only for d < 400 (and every negative d, and NaN): the
interpolated orientation gth is <= 360.0001 whenever the orientation planes hold values in [0, 360] (the library
checks that per keyframe slot at upload and otherwise runs the first-generation loop), and c >= 0.
Run: python tools/verify_gate_algebra2.py   (a few minutes of numpy)."""
import numpy as np

f = np.float32
T2 = f(-45) + f(2.0 ** -16)
assert float(T2).hex() == "-0x1.67fff80000000p+5"
bad2 = bad3 = 0
CH = 1 << 24
with np.errstate(invalid="ignore", over="ignore"):
    for c in range(1 << 8):
        d = (np.arange(CH, dtype=np.uint64) + (c << 24)).astype(np.uint32).view(np.float32)
        dom = ~(d >= f(400))                       # the claimed domain: d < 400 or NaN
        a = np.where(d >= f(360), d - f(360), d)
        a = np.where(a < f(0), a + f(360), a)
        r = np.where(a > f(180), f(360) - a, a)
        r = np.where(r > f(90), f(180) - r, r)
        ref2 = r >= f(80)
        r3 = np.where(a > f(180), f(360) - a, a)
        ref3 = r3 >= f(45)
        a2 = np.where(d < f(0), d + f(360), d)
        new2 = np.abs(np.abs(a2 - f(180)) - f(90)) <= f(10)
        new3 = ((d >= f(45)) | (d <= T2)) & (np.abs(d) <= f(315))
        m2 = (ref2 != new2) & dom
        m3 = (ref3 != new3) & dom
        for name, m in (("gate2", m2), ("gate3", m3)):
            if m.any():
                v = d[m]
                print(name, "differs for", int(m.sum()), "values, e.g.", v[:4], "range", np.nanmin(v), np.nanmax(v),
                      "nan" if np.isnan(v).any() else "")
        bad2 += int(m2.sum()); bad3 += int(m3.sum())
print("gate2 mismatches:", bad2, " gate3 mismatches:", bad3)
