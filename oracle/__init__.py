"""oracle/ -- TEST INFRASTRUCTURE, not product code.

cobs_oracle.c / oracle.py: plain-C restatement of the reference's query algorithm.
ref_shim.cpp / ref.py / _ref/: the unmodified reference, compiled from /root/reference.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
