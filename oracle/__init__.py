"""ORACLE package — CPU restatement of the reference hot path, used only as the
checker by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
