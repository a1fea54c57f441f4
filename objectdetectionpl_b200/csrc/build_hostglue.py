"""Compile hostglue.cpp (pybind11 + libtorch, host only) into the in-tree module the Makefile names.  usage: build_hostglue.py OUT"""
import os
import subprocess
import sys
import sysconfig
import warnings

warnings.filterwarnings("ignore")
import torch
from torch.utils.cpp_extension import include_paths

here = os.path.dirname(os.path.abspath(__file__))
lib = os.path.join(os.path.dirname(torch.__file__), "lib")
cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_hostglue", "-DTORCH_API_INCLUDE_EXTENSION_H",
       "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), os.path.join(here, "hostglue.cpp"), "-o", sys.argv[1]]
cmd += ["-I" + p for p in include_paths() if os.path.isdir(p)] + ["-I" + sysconfig.get_paths()["include"]]
cmd += ["-L" + lib, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-Wl,-rpath," + lib]
sys.exit(subprocess.call(cmd))
