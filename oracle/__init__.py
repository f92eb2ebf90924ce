"""TEST INFRASTRUCTURE — see oracle.cpp. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
