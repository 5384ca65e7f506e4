"""phi.physics.diffuse of the test double (not wrapped by the facade)."""
