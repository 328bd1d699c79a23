"""A plain-C program (examples/c_abi_join.c) compiled against include/dfgpu.h and linked with libdfgpu.so: the header is valid C99,
the library needs nothing but the C runtime at link time, and — on a GPU — the C consumer reproduces the reference's join_inner_one
snapshot in the reference's order.  Without a GPU it must exit with the "no CPU fallback" code."""
import os
import subprocess

import pytest

from datafusion_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="c_abi_join"):
    exe = str(tmp_path / name)
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-Wall", "-Wextra", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", name + ".c"),
                    "-L", libdir, "-ldfgpu", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


@pytest.mark.skipif(capi.load_library().dfgpu_device_count() > 0, reason="a GPU is present")
def test_c_consumer_compiles_links_and_refuses_to_run_without_a_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr
    r = subprocess.run([_build(tmp_path, "c_abi_q3_pipeline")], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_consumer_reproduces_join_inner_one(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split(":")[1].split() for l in r.stdout.splitlines() if l.startswith("row ")]
    assert rows == [["1", "4", "7", "10", "4", "70"], ["2", "5", "8", "20", "5", "80"], ["3", "5", "9", "20", "5", "80"]]


@pytest.mark.gpu
def test_c_consumer_runs_the_fused_q3_plan_with_decimal_money(tmp_path):
    """examples/c_abi_q3_pipeline.c: three fused pipelines (dfgpu_lookup / dfgpu_pipeline) with Decimal128(15,2) money, checked inside
    the program against a nested-loop evaluation of the same tables"""
    r = subprocess.run([_build(tmp_path, "c_abi_q3_pipeline")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = sorted(l.split(":")[1].split() for l in r.stdout.splitlines() if l.startswith("row:"))
    assert rows == [["100", "9000", "0", "9409499.9906"], ["103", "9150", "0", "48323.8100"], ["104", "9203", "1", "26367.3257"]]
    assert "3 groups, 3 expected, equal" in r.stdout
