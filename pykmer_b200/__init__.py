"""pykmer_b200 -- B200-native (sm_100a) indexer and merger hot paths of sauloal/pykmer.

    pykmer_b200.tools     Header / Timer / gen_checksum   (file formats, metadata)
    pykmer_b200.fasta     FASTA text -> cleaned byte stream (host ingest)
    pykmer_b200.indexer   create_fasta_index / main         (.kin, .kin.json)
    pykmer_b200.merger    merge / calculate_distance / main (.kma, .kma.json)
    pykmer_b200.device    Python handles over the C ABI (include/pykmer_b200.h)
    pykmer_b200.synth     synthetic inputs for tests and the benchmark

The compute lives in libpykmer_b200.so (pykmer_b200/csrc, built by
`python -m pykmer_b200.build`).  There is no CPU fallback.
"""
__version__ = "0.1.0"
