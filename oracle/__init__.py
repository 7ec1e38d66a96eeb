"""CPU oracle for the gate-application path — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
