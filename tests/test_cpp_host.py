"""Builds and runs the C++ host-mirror test (tests/cpp/test_host_mirror.cpp): the mirror of the reference's
ChunkGenerator / GenerationParams / generate_mesh interface written in C++ over the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def build():
    import oracle  # builds liboracle.so (checker) as a side effect of lib()
    oracle.lib()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "herbolution_b200")
    odir = os.path.join(ROOT, "oracle", "_build")
    deps = [SRC, os.path.join(libdir, "host", "herbolution_host.hpp"), os.path.join(ROOT, "include", "herbgpu.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", SRC, "-o", EXE, f"-L{libdir}", "-lherbgpu",
           f"-L{odir}", "-loracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}"]
    subprocess.run(cmd, check=True)
    return EXE


def test_cpp_mirror_compiles_and_cpu_checks():
    exe = build()
    out = subprocess.run([exe, "--cpu-only"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_parity_on_gpu():
    exe = build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "byte-exact" in out.stdout


def test_header_is_plain_c():
    """include/herbgpu.h compiles as C11 (-pedantic) and the library links from C; argument checks need no GPU."""
    src = os.path.join(ROOT, "tests", "c", "test_header.c")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_header_c")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(ROOT, "herbolution_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", src, "-o", exe, f"-L{libdir}", "-lherbgpu",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def build_multi():
    src = os.path.join(ROOT, "tests", "cpp", "test_multi.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_multi")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(ROOT, "herbolution_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", src, "-o", exe, f"-L{libdir}", "-lherbgpu",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_multi_driver_compiles():
    build_multi()


@pytest.mark.gpu
def test_multi_gpu_single_process_cpp():
    """hb_multi from C++: one process, all visible GPUs, byte-exact against one device.  On a 1-GPU box the same
    program runs with a one-device handle (the >= 2 GPU run is in profiles/ and in the round-end multi-GPU tier)."""
    import torch
    exe = build_multi()
    args = [exe] if torch.cuda.device_count() >= 2 else [exe, "--allow-one"]
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "byte-exact" in out.stdout
