"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the SSG pseudo-label hot path (reference files cited per function) plus a shim
that imports the unmodified reference from /root/reference when it is present (this container only).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from here.  The product path (self-similarity-grouping_b200/) never does.

Parity status: the reference ships NO tests or golden vectors, so the oracle is
pinned against outputs of the reference itself executed in the build container; the vectors and the
script that made them are committed under tests/golden/ (oracle/make_goldens.py).
"""
