"""TEST INFRASTRUCTURE: stand-ins for the reference's absent third-party imports."""
