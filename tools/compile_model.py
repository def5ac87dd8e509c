#!/usr/bin/env python3
"""MJCF -> .cmodel with the PRODUCT's compiler (csrc/mjcf.cpp), via a tiny host tool built on demand.
usage: compile_model.py in.xml out.cmodel"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'csrc')
TOOL = os.path.join(REPO, 'tools', '_build', 'compile_model')
MAIN = r'''
#include "model.h"
#include <cstdio>
int main(int argc, char **argv) {
  cassie::HostModel m; std::string err;
  if (argc < 3 || !cassie::load_model_any(argv[1], m, err)) { fprintf(stderr, "compile_model: %s\n", err.c_str()); return 1; }
  return cassie::save_cmodel(m, argv[2]) ? 0 : 2;
}
'''


def main():
    os.makedirs(os.path.dirname(TOOL), exist_ok=True)
    src = os.path.join(os.path.dirname(TOOL), 'main.cpp')
    if not os.path.exists(TOOL) or os.path.getmtime(TOOL) < os.path.getmtime(os.path.join(CSRC, 'mjcf.cpp')):
        open(src, 'w').write(MAIN)
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-I', CSRC, src, os.path.join(CSRC, 'mjcf.cpp'), '-o', TOOL])
    subprocess.check_call([TOOL, sys.argv[1], sys.argv[2]])


if __name__ == '__main__':
    main()
