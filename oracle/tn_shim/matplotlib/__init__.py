"""TEST INFRASTRUCTURE ONLY: import stub (helpers.py:17-18)."""
