"""Copies the reference's shipped fixtures that the tests use into tests/golden/.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
These are data files of the reference (profiles, target sequences and the expected outputs
its tutorial documents), not source code.
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [
    "tutorial/AMP_N.bhmm", "tutorial/target-AMP_N.fa", "tutorial/AMP_N-fs.out", "tutorial/AMP_N-fs.tbl",
    "tutorial/PTH2.bhmm", "tutorial/target-PTH2.fa", "tutorial/PTH2.out", "tutorial/PTH2.tbl", "tutorial/PTH2-cigar.tbl",
    "tutorial/tRNA-synthetases.bhmm", "tutorial/PTHR37536.bhmm", "tutorial/AMP_N.out",
    "tutorial/MET-ct4.bhmm", "tutorial/target-MET.fa", "tutorial/MET-ct4.out",
    "tutorial/AMP_N-frameline.out",
    "tutorial/tRNA-proteins.bhmm",      # 12 models whose FS3/FS5 STATS lines one bathconvert run produced: calibration golden values
    "testsuite/2OG-FeII_Oxy_3.bhmm", "testsuite/2OG-FeII_Oxy_3-nt-fs.fa", "testsuite/2OG-FeII_Oxy_3-nt.fa",
]
for f in FILES:
    shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, os.path.basename(f)))
    print("copied", f)
