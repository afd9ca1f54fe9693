#!/usr/bin/env python3
"""Drop-in for the reference's gzireader.py (gzireader.py:21-41): print the block index
`bgzip -i` / `python -m pykmer_b200.bgzf -i` leaves beside a .bgz file.

    gzireader.py <file>.bgz.gzi
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from pykmer_b200.bgzf import print_index  # noqa: E402


def main():
    print_index(sys.argv[1])


if __name__ == "__main__":
    main()
