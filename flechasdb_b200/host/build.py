"""Compiles the C++ host mirror's examples against libflechasdb_b200.so (g++, in-tree)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
EXE = os.path.join(HERE, "example_build_query")
EXE_STORED = os.path.join(HERE, "example_stored")
EXE_TOOL = os.path.join(HERE, "stored_tool")


def _compile(exe, src, extra=()):
    deps = [src, os.path.join(HERE, "flechasdb.hpp"), os.path.join(HERE, "flechasdb_stored.hpp"),
            os.path.join(PKG, "libflechasdb_b200.so"), os.path.join(PKG, "..", "include", "flechasdb_b200.h")]
    if os.path.exists(exe) and all(os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
        return exe
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, src, "-L" + PKG, "-lflechasdb_b200",
                    "-Wl,-rpath," + PKG] + list(extra), check=True)
    return exe


def build():
    _compile(EXE, os.path.join(HERE, "example_build_query.cpp"))
    _compile(EXE_STORED, os.path.join(HERE, "example_stored.cpp"), ["-lz"])
    _compile(EXE_TOOL, os.path.join(HERE, "stored_tool.cpp"), ["-lz"])
    return EXE


if __name__ == "__main__":
    print(build())
