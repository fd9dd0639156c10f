"""kid_b200: B200-native Thompson microphysics step behind the KiD interface."""
