"""The JNI shim is source, not prose: it must parse as C++ against a minimal stand-in jni.h (no JDK in this image) and
reference every entry point include/bsgpu.h declares; the Java class must declare a native method for every JNI export."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_parses_against_stub_jni():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stub_jni"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "bs_jni.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_covers_every_entry_point_and_java_declares_every_export():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bsgpu.h")).read(), flags=re.S)
    api = set(re.findall(r"\b(bs_[a-z0-9_]+)\s*\(", hdr))
    shim = open(os.path.join(ROOT, "jni", "bs_jni.cpp")).read()
    missing = sorted(f for f in api if not re.search(r"\b" + f + r"\s*\(", shim))
    assert not missing, f"bs_jni.cpp never calls {missing}"
    exports = set(re.findall(r"^JF\(\w+, (\w+)\)", shim, flags=re.M))
    java = open(os.path.join(ROOT, "jni", "java", "net", "preibisch", "bigstitcher", "spark", "gpu", "BsNative.java")).read()
    natives = set(re.findall(r"public static native [\w\[\]<>., ]+? (\w+)\(", java))
    assert exports == natives, (sorted(exports - natives), sorted(natives - exports))
