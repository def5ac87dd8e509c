#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/ from the reference checkout.
Runs ONLY in the build container (needs /root/reference); the GPU box uses the committed files.

  *.omodel   oracle model tables: oracle/mjcf_compile.py applied to /root/reference/model/<name>.xml
             (derived numeric tables, not a copy of the XML)
  agility_vectors.npz   input/output pairs of the reference's closed Agility blocks
             (pd_input_step, cassie_core_sim_step from src/libagilitycassie.a via oracle/_ref/liboracle_ref.so)
"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import mjcf_compile as mc  # noqa: E402

REF = os.environ.get('CASSIE_REFERENCE', '/root/reference')
MODELS = ['cassie', 'cassie_hfield', 'cassie_tray_box', 'cassie_no_grav']


def main():
    for name in MODELS:
        m = mc.compile_mjcf(os.path.join(REF, 'model', name + '.xml'))
        mc.write_omodel(m, os.path.join(HERE, name + '.omodel'))
        print('wrote', name + '.omodel')


if __name__ == '__main__':
    main()
