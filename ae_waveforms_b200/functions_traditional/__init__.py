"""Drop-in functions for the reference's ``functions_traditional`` package."""
