#!/usr/bin/env python3
"""Drop-in for the reference CLI:  merger.py <Project_Name> <kin> <kin> [<kin> ...] [--min-count ..]"""
from pykmer_b200.merger import main

if __name__ == "__main__":
    main()
