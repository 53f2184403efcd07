"""TEST INFRASTRUCTURE ONLY: CPU oracles for the sumcheck path (see pymodel.py and zkref.c headers)."""
