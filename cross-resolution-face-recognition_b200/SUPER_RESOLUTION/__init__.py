"""Mirror of the reference's SUPER_RESOLUTION package (the newer FSRNet variant), native ops only."""
