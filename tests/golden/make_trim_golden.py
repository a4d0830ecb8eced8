"""Generates tests/golden/trim_golden.json by running the reference's silence trim (mars5/trim.py, the librosa port that
inference.py:305 applies to the vocoder output) in the build container:  python tests/golden/make_trim_golden.py

mars5/trim.py itself does not run under numpy 2 (as_strided calls `np.array(x, copy=False)`, trim.py:546, which numpy 2
turned into an error; numpy's message: "replace it with np.asarray(obj) ... no behavior change in NumPy 1.x").  The
reference file is left untouched: this script hands the module a numpy proxy whose `array(..., copy=False)` does what
numpy 1.x did, and everything else is the reference's own arithmetic.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mars5.trim as ref_trim  # noqa: E402
from tests.golden.inputs import trim_cases  # noqa: E402


class _Numpy1Array:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(x, copy=True, subok=False, **kw):
        return np.asarray(x) if copy is False else np.array(x, copy=copy, subok=subok, **kw)


ref_trim.np = _Numpy1Array()
out = {}
for name, y in trim_cases():
    for top_db in (27, 60):
        yt, idx = ref_trim.trim(y, top_db=top_db)
        out[f"{name}@{top_db}"] = {"n": int(y.numel()), "start": int(idx[0]), "end": int(idx[1]), "len": int(yt.numel())}
        assert yt.numel() == int(idx[1]) - int(idx[0])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trim_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: (v["start"], v["end"]) for k, v in out.items()})
