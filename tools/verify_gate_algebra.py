"""Exhaustive check (all 2^32 float bit patterns) that the short forms of the two direction gates used by the
scan kernel decide exactly like the reference's fold sequences (ProbabilityMapping.cc:795-809), for any float
`a` = the angle difference after the two wrap steps  `if (a >= 360) a -= 360; if (a < 0) a += 360;`.

  gate 2 (:797-798): if (a > 180) a = 360 - a;  if (a > 90) a = 180 - a;  reject iff a >= lambdaL(80)
      short form: reject iff  | |a - 180| - 90 | <= 10        (NaN: not rejected, like the reference)
  gate 3 (:808-809): if (a > 180) a = 360 - a;                          reject iff a >= lambdaTheta(45)
      short form: reject iff  a >= 45 && a <= 315
Run: python tools/verify_gate_algebra.py   (about 5 minutes of numpy)."""
import numpy as np

f = np.float32
bad2 = bad3 = 0
CH = 1 << 24
with np.errstate(invalid="ignore", over="ignore"):
    for c in range(1 << 8):
        a = (np.arange(CH, dtype=np.uint64) + (c << 24)).astype(np.uint32).view(np.float32)
        # reference gate 2
        r = np.where(a > f(180), f(360) - a, a)
        r = np.where(r > f(90), f(180) - r, r)
        ref2 = r >= f(80)
        new2 = np.abs(np.abs(a - f(180)) - f(90)) <= f(10)
        # reference gate 3
        r3 = np.where(a > f(180), f(360) - a, a)
        ref3 = r3 >= f(45)
        new3 = (a >= f(45)) & (a <= f(315))
        d2 = ref2 != new2
        d3 = ref3 != new3
        if d2.any() or d3.any():
            for name, d in (("gate2", d2), ("gate3", d3)):
                if d.any():
                    v = a[d]
                    print(name, "differs for", int(d.sum()), "values, e.g.", v[:4], "range", np.nanmin(v), np.nanmax(v),
                          "nan" if np.isnan(v).any() else "")
        bad2 += int(d2.sum()); bad3 += int(d3.sum())
print("gate2 mismatches:", bad2, " gate3 mismatches:", bad3)
