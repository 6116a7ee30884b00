#!/usr/bin/env python
"""Generate tests/golden/ref_flo.npz: a small .flo file (bytes) and the (u, v) the REAL reference's
utils/flow_io.py::flow_read returns for it.  Run in the authoring container only:

    python tests/golden/make_golden_flo.py

The file is written by this script (the reference's own flow_write fails on Python 3: it writes a ``str`` tag to a
binary file, flow_io.py:78), with payload values that exercise a de-interleave: distinct per element, signed zeros,
infinities, a NaN, Sintel's 1e9 'unknown' marker, odd width and height.
"""
import importlib.util
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    spec = importlib.util.spec_from_file_location("ref_flow_io", os.path.join(REF, "utils", "flow_io.py"))
    fio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fio)
    rs = np.random.RandomState(4242)
    h, w = 37, 53
    u = (rs.standard_normal((h, w)) * 20).astype(np.float32)
    v = (rs.standard_normal((h, w)) * 20).astype(np.float32)
    u[0, 0], v[0, 0] = 0.0, -0.0
    u[1, 2], v[1, 2] = np.inf, -np.inf
    u[2, 3] = np.nan
    u[3, 4], v[3, 4] = 1e9, 1e9
    payload = np.empty((h, 2 * w), np.float32)
    payload[:, 0::2] = u
    payload[:, 1::2] = v
    raw = b"PIEH" + np.array([w, h], np.int32).tobytes() + payload.tobytes()
    with tempfile.NamedTemporaryFile(suffix=".flo", delete=False) as f:
        f.write(raw)
        name = f.name
    ru, rv = fio.flow_read(name)
    os.unlink(name)
    assert np.array_equal(ru.view(np.uint32), u.view(np.uint32)) and np.array_equal(rv.view(np.uint32), v.view(np.uint32))
    np.savez_compressed(os.path.join(HERE, "ref_flo.npz"), flo=np.frombuffer(raw, np.uint8), u=np.ascontiguousarray(ru),
                        v=np.ascontiguousarray(rv))
    print("ref_flo.npz:", len(raw), "bytes of .flo,", ru.shape)


if __name__ == "__main__":
    main()
