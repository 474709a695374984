"""TEST INFRASTRUCTURE ONLY: import stub."""
