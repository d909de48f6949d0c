"""INI configuration loader with the reference's contract (config.py:9-12 of the reference):
``load_conf_info(path) -> configparser.ConfigParser``; a missing file yields an empty parser and
missing keys raise configparser errors at the point of use."""
import configparser


def load_conf_info(config_file):
    parser = configparser.ConfigParser()
    parser.read(config_file)
    return parser


def section_with(config, option, preferred=("testing", "inference")):
    """The reference's InferenceEngine reads ('testing','checkpoint_filepath') although the
    shipped infer cfg names its section [inference]: accept either."""
    for sec in preferred:
        if config.has_section(sec) and config.has_option(sec, option):
            return sec
    # fall through to the reference's behaviour: raise NoSectionError/NoOptionError on 'testing'
    config.get(preferred[0], option)
    return preferred[0]
