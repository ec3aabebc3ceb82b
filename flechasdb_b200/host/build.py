"""Compiles the C++ host mirror's example against libflechasdb_b200.so (g++, in-tree)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
EXE = os.path.join(HERE, "example_build_query")


def build():
    src = os.path.join(HERE, "example_build_query.cpp")
    deps = [src, os.path.join(HERE, "flechasdb.hpp"), os.path.join(PKG, "libflechasdb_b200.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", EXE, src, "-L" + PKG, "-lflechasdb_b200",
                    "-Wl,-rpath," + PKG], check=True)
    return EXE


if __name__ == "__main__":
    print(build())
