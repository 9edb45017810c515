"""The JNI glue a maintainer adds on the reference side (shim/jni/sgr_jni.c) cannot be built here — no JDK — but it can be
type-checked against include/sgr.h with a stand-in <jni.h> (tests/mock_jni): a renamed C entry point, a wrong argument
count or a missing native would otherwise only surface on the maintainer's machine. Also checks that every @native of
shim/scala/surge/gpu/Native.scala has its Java_surge_gpu_Native_00024_<name> definition in the C file."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jni_glue_type_checks_against_the_c_abi():
    cmd = ["gcc", "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
           "-I", os.path.join(ROOT, "tests", "mock_jni"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "shim", "jni", "sgr_jni.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_every_scala_native_has_its_c_definition():
    scala = open(os.path.join(ROOT, "shim", "scala", "surge", "gpu", "Native.scala")).read()
    c_src = open(os.path.join(ROOT, "shim", "jni", "sgr_jni.c")).read()
    natives = re.findall(r"@native\s+def\s+(\w+)", scala)
    assert len(natives) >= 20
    defined = set(re.findall(r"Java_surge_gpu_Native_00024_(\w+)\s*\(", c_src))
    assert set(natives) <= defined, sorted(set(natives) - defined)
    assert defined <= set(natives), sorted(defined - set(natives))


def test_the_c_abi_header_is_self_contained_in_c_and_cxx(tmp_path):
    src = tmp_path / "only_sgr.c"
    src.write_text('#include "sgr.h"\nint main(void) { return SGR_ABI_VERSION - 1; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
