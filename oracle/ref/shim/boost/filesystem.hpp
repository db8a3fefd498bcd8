// ORACLE / TEST INFRASTRUCTURE ONLY: nothing of boost::filesystem is used on the path that is compiled.
#ifndef STOMP_B200_ORACLE_BOOST_FS_SHIM
#define STOMP_B200_ORACLE_BOOST_FS_SHIM
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem { inline bool create_directories(const std::string& p) { return ::mkdir(p.c_str(), 0755) == 0; } inline bool create_directory(const std::string& p) { return ::mkdir(p.c_str(), 0755) == 0; } inline bool exists(const std::string& p) { struct stat s; return ::stat(p.c_str(), &s) == 0; } } }
#endif
