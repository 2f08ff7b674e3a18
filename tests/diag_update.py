"""Diagnostic (not a test): print every CUDA-vs-oracle error of the phased (teacher-forced)
update for the given scenarios without asserting.
usage: python tests/diag_update.py [scenario ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_update_parity_gpu as T


def main():
    for name in (sys.argv[1:] or list(T.ALL)):
        print(T.run_phased(name, check=False).text())


if __name__ == '__main__':
    main()
