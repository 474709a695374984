"""TEST INFRASTRUCTURE ONLY (type-annotation stub, node_array.py:21)."""
