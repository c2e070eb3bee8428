"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/crane_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package. Parity status: UNPINNED by the
reference (no scheduler tests upstream, reference not buildable here).
"""
