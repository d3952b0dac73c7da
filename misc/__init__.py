"""Import-path compatible facade of the reference's ``misc`` package."""
