#!/usr/bin/env python3
"""Drop-in for the reference CLI:  calculate_distance.py <project.MIN-MAX.kma>"""
from pykmer_b200.distance import main

if __name__ == "__main__":
    main()
