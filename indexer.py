#!/usr/bin/env python3
"""Drop-in for the reference CLI:  indexer.py <input_file> <sample_name> <kmer_len>"""
from pykmer_b200.indexer import main

if __name__ == "__main__":
    main()
