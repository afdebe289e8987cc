"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference path + harness around the real reference.
Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs — never by mixermdm_b200."""
