/* TEST INFRASTRUCTURE ONLY - stand-in for <boost/algorithm/string.hpp> (absent in this image) so
 * that the reference's source/PathStore.cpp compiles where it lies: it only calls boost::trim. */
#pragma once
#include <string>
namespace boost
{
	inline void trim(std::string& text)
	{
		const char* blanks = " \t\r\n\v\f";
		const size_t first = text.find_first_not_of(blanks);
		const size_t last = text.find_last_not_of(blanks);
		text = (first == std::string::npos) ? std::string() : text.substr(first, last - first + 1);
	}
}
