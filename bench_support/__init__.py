"""Benchmark support: synthetic workload generators (not part of the codec)."""
