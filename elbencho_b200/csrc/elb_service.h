/*
 * Distributed mode (--service / --hosts): the reference's HTTP + JSON wire format
 * ("next" row f3 of SURVEY.md §8): endpoints and key names of source/Common.h:84,200-269,
 * service side source/HTTPServiceSWS.cpp, master side source/workers/RemoteWorker.cpp, config
 * transfer source/ProgArgs.cpp:3562-3863, status/result trees source/Statistics.cpp:1350-1405,
 * 2728-2804, histogram (de)serialisation source/LatencyHistogram.cpp:68-97.
 *
 * Dependency-free: a small ordered JSON tree (boost::property_tree semantics: every leaf is a
 * string, duplicate keys allowed and kept in order), a poll()-based HTTP/1.1 server and a
 * blocking HTTP client.
 */
#ifndef ELB_SERVICE_H_
#define ELB_SERVICE_H_

#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "elb_cli.h"

#define ELB_HTTP_PROTOCOLVERSION "3.1.1" /* source/Common.h:84 */

namespace elb
{

/* boost::property_tree::ptree stand-in */
class JsonTree
{
	public:
		typedef std::vector<std::pair<std::string, JsonTree> > ChildVec;

		JsonTree() {}
		explicit JsonTree(const std::string& value) : value(value) {}

		// put: replace the first child of that (dotted) path or append; add: always append
		void put(const std::string& path, const std::string& newValue);
		void put(const std::string& path, uint64_t newValue) { put(path, std::to_string(newValue) ); }
		void putBool(const std::string& path, bool newValue) { put(path, newValue ? "true" : "false"); }
		void add(const std::string& path, const std::string& newValue);
		void add(const std::string& path, uint64_t newValue) { add(path, std::to_string(newValue) ); }

		bool has(const std::string& path) const { return find(path) != NULL; }
		std::string getStr(const std::string& path) const; // @throw ProgError if missing
		std::string getStr(const std::string& path, const std::string& defaultValue) const;
		uint64_t getU64(const std::string& path) const;
		uint64_t getU64(const std::string& path, uint64_t defaultValue) const;
		bool getBool(const std::string& path) const;
		bool getBool(const std::string& path, bool defaultValue) const;
		const JsonTree* find(const std::string& path) const;

		const ChildVec& getChildren() const { return children; }
		const std::string& getValue() const { return value; }

		std::string toJSON(bool pretty = true) const;
		static JsonTree parse(const std::string& text); // @throw ProgError

	private:
		std::string value;
		ChildVec children;

		JsonTree* findOrCreate(const std::string& path, bool alwaysAppendLeaf);
		void write(std::string& out, bool pretty, int indent) const;
};

struct HttpRequest
{
	std::string method;
	std::string path;
	std::map<std::string, std::string> query;
	std::string body;
	std::string remoteAddr;
};

struct HttpResponse
{
	int statusCode{200};
	std::string body;
};

/* blocking client: one request per connection ("Connection: close") */
HttpResponse httpRequest(const std::string& host, unsigned short port, const std::string& method,
	const std::string& pathAndQuery, const std::string& body, int timeoutSecs);

std::string urlEncode(const std::string& raw);

int serviceMain(ProgArgs& progArgs);
int masterMain(ProgArgs& progArgs);
int masterInterruptOrQuitServices(ProgArgs& progArgs);

} // namespace elb

#endif /* ELB_SERVICE_H_ */
