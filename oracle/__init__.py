"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the per-frame CRT effect chain of PythonCRT
(`/root/reference/crt_filter.py:531-699`, `:702-861`, `:1086-1098`) used as the
parity checker for the CUDA path and as the timed CPU baseline in `bench.py`.

Nothing in the product package (`pythoncrt_b200/`) imports this directory: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may.  The product path fails loudly when the CUDA
library is missing; it never falls back to this code.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against the reference itself, imported unmodified in the
build container by `oracle/ref_loader.py`, and against fixtures generated from
it by `oracle/make_golden.py` (committed under `tests/golden/`).  The stencil /
gather arithmetic of the reference lives in OpenCV and numpy (requirements.txt
gives lower bounds only); the fixtures record the versions they were made with
(numpy 2.3.5, opencv-python-headless 4.13.0).
"""
