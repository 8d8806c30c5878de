"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's `Stereo.get_depth` hot path (calibrating/stereo_camera.py:492-533,
calibrating/stereo_matching.py:22-70, calibrating/utils.py:173-200).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package; nothing under `calibrating_b200/` does.
"""
