"""Helpers mirroring the reference's ``collectivecrossing.utils`` package."""
