/* TEST INFRASTRUCTURE -- stands in for the autoconf-generated config.h of the
 * reference (configure.ac); every reference source includes it under
 * HAVE_CONFIG_H (e.g. ModPlugin.h:28-30). Optional outputs stay disabled. */
#define PACKAGE "odr-dabmod"
#define PACKAGE_NAME "odr-dabmod"
#define PACKAGE_VERSION "3.0.1"
#define VERSION "3.0.1"
#define HAVE_PRCTL 1
#define HAVE_NETINET_IN_H 1
