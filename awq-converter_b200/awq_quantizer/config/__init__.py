"""Configuration package."""
