"""CPU stress of the worker pool behind vit_run's pageable-buffer path (csrc/vit_stage_pool.h, host-only): built under
ThreadSanitizer and hammered with back-to-back jobs of varying size -- every item exactly once, inside its own call, no
data race reported.  (Its first version published the job's count and function as plain members; TSan flags that version.)"""
import os
import subprocess

import pytest

from vit_testlib import ROOT

GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SRC = os.path.join(ROOT, "tests", "host", "stage_pool_stress.cpp")


def _build(tmp_path, flags, name):
    exe = tmp_path / name
    r = subprocess.run([GXX, "-O1", "-g", "-std=c++17", "-pthread"] + flags + ["-o", str(exe), SRC], capture_output=True, text=True)
    return (str(exe), r)


def test_stage_pool_stress_native(tmp_path):
    exe, r = _build(tmp_path, [], "stress")
    assert r.returncode == 0, r.stderr[-800:]
    out = subprocess.run([exe, "8", "2"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and " errors 0 " in out.stdout, out.stdout + out.stderr[-400:]


def test_stage_pool_is_race_free_under_thread_sanitizer(tmp_path):
    exe, r = _build(tmp_path, ["-fsanitize=thread"], "stress_tsan")
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    out = subprocess.run([exe, "6", "3"], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in out.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container: " + out.stderr[-200:])
    assert "WARNING: ThreadSanitizer" not in out.stderr, out.stderr[:3000]
    assert out.returncode == 0 and " errors 0 " in out.stdout, out.stdout
