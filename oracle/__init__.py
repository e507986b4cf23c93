"""CPU oracle for the NeRF ray-march hot path. TEST INFRASTRUCTURE ONLY: see ray_march.py."""
