"""ORACLE package: CPU restatements of the reference's TTS tail.

Test infrastructure only.  Nothing under infernos_b200/ imports it; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg do.
"""
