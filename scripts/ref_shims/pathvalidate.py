"""Stand-in for pathvalidate.  Golden-generation infrastructure only (see README.md)."""


class ValidationError(Exception):
    pass


def validate_filepath(p, platform=None):
    if not isinstance(p, str) or '\n' in p or '\0' in p or len(p) > 1024:
        raise ValidationError(p)
