"""CPU oracle for the BAREFOOT hot path.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU arms, never by the product."""
