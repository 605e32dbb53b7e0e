"""Host-side logic of bench.py that needs no GPU: the work list of the reference arm, the shared config object and the timeline
writer (fed with a fake context)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_cell_frame_task_list_matches_the_frame_structure():
    """One cfg2 cell-frame (TDD DDDSU @30 kHz): 12 DL slots x (PDSCH + DM-RS) precoding calls, 4 CSI-RS occasions x 8 UEs,
    5 SRS occasions x 4 UEs, one sensing pass -- the same counts CommWorkload.step issues on the GPU leg."""
    b = _bench()
    tasks = b.cell_frame_tasks()
    kinds = [t[0] for t in tasks]
    assert kinds.count("prg") == 24 and kinds.count("csi") == 32 and kinds.count("srs") == 20 and kinds.count("sense") == 1
    assert len(tasks) == 77
    assert sorted({t[1] for t in tasks if t[0] == "csi"}) == [2, 7, 12, 17]
    assert sorted({t[1] for t in tasks if t[0] == "srs"}) == [3, 4, 11, 12, 19]


def test_config_object_is_shared_by_both_arms():
    b = _bench()
    c = b.bench_config(4, 10)
    assert c == b.bench_config(4, 10) and c["workload"].startswith("cfg2") and c["frames_per_step"] == 10
    assert "model" not in c and c["stages"][0] == "echo_demod"


def test_reference_arm_precoding_task_runs_on_the_oracle():
    b = _bench()
    secs, fft = b.run_task(("prg", 0, 1), 3)      # the DM-RS allocation: the cheapest task
    assert secs > 0 and fft == 0.0


class _FakeCtx:
    def __init__(self, rec):
        self.rec = rec

    def profile_timeline(self, base):
        return self.rec


def test_timeline_writer_reports_busy_time_and_gaps(tmp_path):
    b = _bench()
    # two frames of 1 ms: [pmi 0-0.6][gap 0.1][cdl 0.7-0.9] and the sensing context's echo group inside the second frame
    a = _FakeCtx([("pmi_sinr", 0.0, 0.6), ("cdl", 0.7, 0.9), ("pmi_sinr", 1.0, 1.6), ("cdl", 1.9, 2.0)])
    s = _FakeCtx([("echo_demod", 1.6, 1.8)])
    path = tmp_path / "tl.txt"
    b.write_timeline(str(path), None, [a, s], 2.0, 2)
    txt = path.read_text().splitlines()
    assert "5 kernel groups over 2 frames" in txt[0] and "1.700 ms inside groups (85.0 %)" in txt[0]
    pairs = {tuple(l.split()[1:4:2]): float(l.split()[4]) for l in txt if l.startswith("#   ")}
    assert np.isclose(pairs[("pmi_sinr", "cdl")], 50.0)         # 0.1 ms over two frames
    assert np.isclose(pairs[("cdl", "pmi_sinr")], 50.0) and np.isclose(pairs[("echo_demod", "cdl")], 50.0)
    body = [l for l in txt if not l.startswith("#")]
    assert [l.split()[-1] for l in body] == ["pmi_sinr", "echo_demod", "cdl"]   # the second frame, in launch order
